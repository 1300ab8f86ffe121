#!/bin/bash
# short bench (no cpu baseline) + cache-preserving launch list of one beam decode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','greedy_captions_per_s')}); print(d['train'] and (d['train']['value'], d['train']['ms_per_step'])); print(d['roofline']['us_per_launch'], d['roofline']['frac'])"; tail -3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_beam_nocc.csv python scripts/profile_step.py beam > /dev/null 2>&1; echo "ncu beam list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_beam_nocc.csv 2>/dev/null | head -12
