#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q --no-header -p no:cacheprovider --timeout 240 tests/test_gpu_kernels.py -m gpu -k gemm 2>&1 | tail -4
timeout 300 python scripts/gemm_trace.py
bash scripts/gpu_quick.sh
