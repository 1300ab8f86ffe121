#!/bin/bash
# compute-sanitizer over the kernels with hand-written synchronisation (mbarrier rings, the last-arriver merge of the attention
# kernel, the TMEM / TMA pipeline of the GEMM, the fused beam step): memcheck, then racecheck (shared-memory hazards).
# Usage: gpurun --timeout 1500 -- scripts/sanitize.sh [tag]      -> gpurun_out/<tag>_sanitize_{memcheck,racecheck}.txt
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
sel='att_step_fwd and (5-1-7 or 3-2-196 or 16-1-196) or beam_advance_equals and 5-3-3 or gemm_tn_matches_fp32 and 300-200-72 or logit_stats_topk and 6-52'
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 120 \
    python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "$sel" > gpurun_out/${tag}_sanitize_${tool}.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/${tag}_sanitize_${tool}.txt | tail -5
done
# the batch-balanced attention kernel (attention_v7.cu: release / acquire flags between CTAs, shared-memory atomics): memcheck
sel7='v7_with_region_masks and (200-2-52 or 150-5-40) or launch_plans and (148-3-196 or 300-3-100) or wide_beam_groups and 80-5-100'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 1 --launch-timeout 120 \
  python -m pytest tests/test_gpu_bench_plans.py -q -x -p no:cacheprovider -k "$sel7" > gpurun_out/${tag}_sanitize_memcheck_v7.txt 2>&1
echo "memcheck v7 exit $?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/${tag}_sanitize_memcheck_v7.txt | tail -5

