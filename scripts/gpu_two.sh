#!/bin/bash
# 2-GPU check of the bench contract (torchrun, NCCL): own arm + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}); print(d['e2e']); print(d['train'])"
tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "ref 2gpu exit $?"; cut -c1-200 gpurun_out/bench_ref_2gpu.json
