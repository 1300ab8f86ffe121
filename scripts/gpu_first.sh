#!/bin/bash
# compact first beam step: kernel + model parity, then the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_oracle.py tests/test_gpu_shapes.py tests/test_gpu_step_api.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_first.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s','gpu_launches')})"
