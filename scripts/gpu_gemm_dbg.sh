#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0; do echo "=== UIC_GEMM_DEBUG=$d"; UIC_GEMM_DEBUG=$d timeout 300 python scripts/gemm_trace.py 2>&1 | grep -E "^---|last mma" | sed -n 3,6p; done
