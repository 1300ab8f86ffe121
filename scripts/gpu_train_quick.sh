#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q --no-header -p no:cacheprovider --timeout 240 tests/test_gpu_train.py tests/test_gpu_sampling.py -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s')}); print(d['train'] and (d['train']['value'], d['train']['ms_per_step']))"; tail -2 gpurun_out/bench.err
