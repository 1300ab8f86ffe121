#!/bin/bash
# ncu --set full capture of the attention-step kernel at the benchmark shape (scripts/att_bench.py, first case).
# Usage: gpurun --timeout 600 -- scripts/gpu_ncu_att.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-att}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:att_step_fwd -s 0 -c 1 -o gpurun_out/ncu_${tag} -f \
  python scripts/att_bench.py > gpurun_out/ncu_${tag}.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/ncu_${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_ncu_raw.csv
