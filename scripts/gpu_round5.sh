#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/grad_errs.jsonl
bash scripts/gpu_tests.sh
for f in gemm att misc golden oracle train sampling; do grep -h -E "^(FAILED|ERROR)" gpurun_out/test_$f.log | head -6; done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s','gpu_launches')}); print(d['e2e']); print(d['train'] and (d['train']['value'], d['train']['ms_per_step'])); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks'])"; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_beam.csv python scripts/profile_step.py beam > /dev/null 2>&1; echo "ncu beam list exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_step.py train > /dev/null 2>&1; echo "ncu train list exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:att_step_fwd -c 1 -o gpurun_out/att_full -f python scripts/profile_step.py beam > gpurun_out/ncu_att.log 2>&1; echo "ncu att exit $?"
