#!/bin/bash
# Run the GPU test-suite in separate processes (a sticky CUDA error in one group must not hide the
# others) and keep the logs under gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, pytest args...
  local name=$1; shift
  timeout 600 python -m pytest -q --no-header -p no:cacheprovider --timeout 180 "$@" > gpurun_out/test_$name.log 2>&1
  echo "== $name: exit $? :: $(tail -1 gpurun_out/test_$name.log)"
}
run gemm tests/test_gpu_kernels.py -m gpu -k gemm
run att tests/test_gpu_kernels.py -m gpu -k att_step
run misc tests/test_gpu_kernels.py -m gpu -k "not gemm and not att_step"
run golden tests/test_gpu_golden.py -m gpu
run oracle tests/test_gpu_oracle.py -m gpu
run train tests/test_gpu_train.py -m gpu
run sampling tests/test_gpu_sampling.py -m gpu
for f in "$@"; do :; done
