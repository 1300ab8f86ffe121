#!/bin/bash
# programmatic dependent launch A/B: parity suites with PDL on, then the bench with PDL on and off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q --no-header -p no:cacheprovider --timeout 240 tests/test_gpu_golden.py tests/test_gpu_oracle.py tests/test_gpu_sampling.py tests/test_gpu_kernels.py -m gpu 2>&1 | tail -5
for pdl in 1 0; do
UIC_PDL=$pdl timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "PDL=$pdl bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$pdl.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s')}); print(d['train'] and (d['train']['value'], d['train']['ms_per_step']))"; tail -2 gpurun_out/bench_pdl$pdl.err
done
