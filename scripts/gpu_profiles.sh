#!/bin/bash
# Reproduces the measurement artefacts under profiles/ (tag = file prefix, e.g. r2):
#   launch lists of ONE beam-3 decode and ONE training step (ncu --metrics gpu__time_duration.sum --clock-control none),
#   ncu --set full raw pages of the attention step kernel and of the GEMM variants on the benchmark path,
#   the default bench.py line and the --impl reference line.
# Usage: gpurun --timeout 2400 -- scripts/gpu_profiles.sh r2
cd "$(dirname "$0")/.."
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err; echo "reference exit $?"
for mode in beam train; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches_${mode}.csv python scripts/profile_step.py $mode > /dev/null 2>&1; echo "ncu $mode list exit $?"
  python scripts/summarize_launches.py $out/${tag}_launches_${mode}.csv > $out/${tag}_launches_${mode}_summary.txt
done
cap() {  # name regex mode skip
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $4 -c 1 \
    -o $out/ncu_$1 -f python scripts/profile_step.py $3 > $out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i $out/ncu_$1.ncu-rep --page raw --csv > $out/${tag}_$1_ncu_raw.csv 2>/dev/null
  rm -f $out/ncu_$1.ncu-rep     # ~40 MB each: over the 64 MiB return limit; the raw page is what gets committed
}
# GEMM launches of one beam decode, in order: 0 att_embed, 1 ctx2att, then per step gates / a2c / statistics (step 0 runs on
# one row per image, so the 768-row shapes start at launch 5)
cap att_step_fwd att_step_fwd beam 5
cap gemm_prologue gemm_bf16_tcgen05_kernel beam 0
cap gemm_gates gemm_bf16_tcgen05_kernel beam 5
cap gemm_stats gemm_bf16_tcgen05_kernel beam 7
python scripts/ncu_summary.py $out/${tag}_*_ncu_raw.csv > $out/${tag}_ncu_summary.txt; cat $out/${tag}_ncu_summary.txt
