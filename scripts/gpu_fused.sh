#!/bin/bash
# fused vocabulary statistics: kernel tests, golden/oracle parity, bench without cpu baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 tests/test_gpu_kernels.py -m gpu -k "logit_stats or greedy_merge or topk or greedy_step or advance" 2>&1 | tail -15
timeout 900 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 tests/test_gpu_oracle.py tests/test_gpu_golden.py -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s','gpu_launches')}); print(d['roofline']['us_per_launch'], d['roofline']['frac'])
for k,v in d['kernel_shares'].items(): print(k, round(v['ms_per_step']*1000/ max(1,v['launches']),1), 'us/launch', round(v['share'],3))"; tail -3 gpurun_out/bench.err
