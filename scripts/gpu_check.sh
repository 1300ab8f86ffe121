#!/bin/bash
# Round-end check as the driver runs it: the whole GPU suite in one process (all failures listed, not -x), smoke(),
# the default bench line.  Usage: gpurun --timeout 1500 -- scripts/gpu_check.sh [tag]
cd "$(dirname "$0")/.."
tag=${1:-check}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -n 40 gpurun_out/${tag}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -n 5 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench.json'))
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'gpu_launches')})
for k in ('e2e', 'roofline', 'cpu_baseline', 'clocks'):
    print(k, d.get(k))
PY
