#!/bin/bash
# one ncu --set full capture of a kernel (regex $1) inside a beam decode or a training step ($2 = beam|greedy|train)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c 1 -o gpurun_out/ncu_$1 -f python scripts/profile_step.py $2 > gpurun_out/ncu_$1.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_$1.log
