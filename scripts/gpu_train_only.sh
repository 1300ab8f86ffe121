#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/grad_errs.jsonl
timeout 600 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 tests/test_gpu_train.py -m gpu > gpurun_out/test_train.log 2>&1
echo "train exit $? :: $(tail -1 gpurun_out/test_train.log)"
grep -h -E "^E .*(Error|assert)" gpurun_out/test_train.log | head -20
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/grad_errs.jsonl')]
for r in rows:
    worst=sorted(r.items(), key=lambda kv:-kv[1])[:4]
    print(worst)
PY
