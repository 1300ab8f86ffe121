#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
grep -h -E "^(FAILED|ERROR)|Error|assert " gpurun_out/test_train.log | head -40
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s','train')}); print(d['e2e'])"; tail -5 gpurun_out/bench.err
