#!/bin/bash
# Generic lease script: runs the given commands on the GPU box with stdout/stderr kept under gpurun_out/<tag>.log
# Usage: gpurun --timeout 900 -- scripts/gpu_run.sh <tag> '<shell command>'
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_out
bash -c "$*" > gpurun_out/${tag}.log 2>&1
echo "exit $?"; tail -n 60 gpurun_out/${tag}.log
