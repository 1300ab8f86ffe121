#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train_nocc.csv python scripts/profile_step.py train > /dev/null 2>&1; echo "ncu train list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_train_nocc.csv 2>/dev/null | head -32
