#!/bin/bash
# diverse beam search: kernel test + model test against the oracle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_oracle.py -q -m gpu -k "diverse or beam" 2>&1 | tail -30
