#!/bin/bash
# final launch lists of round 1 (beam decode + training step) for profiles/r1_pass10_*
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_beam.csv python scripts/profile_step.py beam > /dev/null 2>&1; echo "ncu beam list exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_step.py train > /dev/null 2>&1; echo "ncu train list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_beam.csv > gpurun_out/launches_beam_summary.txt 2>&1; head -14 gpurun_out/launches_beam_summary.txt
python scripts/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.txt 2>&1; head -8 gpurun_out/launches_train_summary.txt
