#!/bin/bash
# ncu --set full captures of the tcgen05 GEMM variants on the benchmark path (tensor-pipe evidence):
#   <256,3,0,0,0> feature prologue (50176 x 512 x 2048), <128,5,0,0,0> per-step gates GEMM (768 x 3072 x 1024),
#   <128,5,0,0,3> statistics GEMM (768 x 10000 x 512), and the time-batched logit GEMM of a training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex mode skip
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $4 -c 1 -o gpurun_out/ncu_$1 -f python scripts/profile_step.py $3 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i gpurun_out/ncu_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_$1.ncu-rep     # ~40 MB each: over the 64 MiB return limit; the raw page is what gets committed
  python - "$1" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/ncu_{name}_raw.csv")))
hdr, vals = rows[0], rows[-1]
want = ("Kernel Name", "Demangled Name", "Function Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size")
for k in want:
    for i, h in enumerate(hdr):
        if h == k or h.startswith(k):
            print(f"  {h}: {vals[i][:90]}")
            break
PY
}
# ncu matches the base name only; the GEMM launches of one beam decode come in this order: 0 att_embed, 1 ctx2att, then per
# step gates / a2c / statistics (step 0 runs on one row per image, so the 768-row shapes start at launch 5)
cap gemm_prologue gemm_bf16_tcgen05_kernel beam 0
cap gemm_gates gemm_bf16_tcgen05_kernel beam 5
cap gemm_stats gemm_bf16_tcgen05_kernel beam 7
