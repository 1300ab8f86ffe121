#!/bin/bash
# launch list with the caches left as the previous kernel left them (closer to the in-graph durations)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_beam_nocc.csv python scripts/profile_step.py beam > /dev/null 2>&1; echo "ncu beam list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_beam_nocc.csv | head -14
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train_nocc.csv python scripts/profile_step.py train > /dev/null 2>&1; echo "ncu train list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_train_nocc.csv | head -24
