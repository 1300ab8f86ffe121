#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 900 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json | cut -c1-600
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_beam.csv python scripts/profile_step.py beam > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:att_step_fwd -c 2 -o gpurun_out/att_full -f python scripts/profile_step.py beam > gpurun_out/ncu_att.log 2>&1; echo "ncu att exit $?"
ls -la gpurun_out | head -30
