#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/grad_errs.jsonl
bash scripts/gpu_tests.sh
for f in att golden oracle train misc; do grep -h -E "^(FAILED|ERROR)" gpurun_out/test_$f.log | head -8; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s')}); print(d['e2e']); print(d['train'] and d['train']['value']); print(d['roofline'])
for k,v in d['kernel_shares'].items(): print(k, v)"; tail -3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:att_step_fwd -c 1 -o gpurun_out/att4_full -f python scripts/profile_step.py beam > gpurun_out/ncu_att.log 2>&1; echo "ncu att exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 2 -c 3 -o gpurun_out/gemm_full -f python scripts/profile_step.py beam > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
