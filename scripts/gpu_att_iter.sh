#!/bin/bash
# attention-kernel iteration: unit tests, timeline of CTA 0, golden/oracle parity, short bench, optional ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 tests/test_gpu_kernels.py -m gpu -k "att_step" 2>&1 | tail -5
timeout 300 python scripts/att_trace.py 2>&1 | tail -20
timeout 600 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 tests/test_gpu_oracle.py tests/test_gpu_golden.py -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s')}); print(d['roofline']['us_per_launch'], d['roofline']['frac'])
for k,v in d['kernel_shares'].items(): print(k, round(v['ms_per_step']*1000/ max(1,v['launches']),1), 'us/launch', round(v['share'],3))"; tail -3 gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:att_step_fwd -c 1 -o gpurun_out/att_quick -f python scripts/profile_step.py beam > gpurun_out/ncu_att.log 2>&1; echo "ncu att exit $?"
fi
