#!/bin/bash
# N-GPU check of the bench contract (torchrun, NCCL): own arm only.  usage: gpu_n.sh N
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1; lscpu | grep -i numa > gpurun_out/numa_${N}gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench ${N}gpu exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}); print(d['e2e']); print(d['train']); print(d.get('greedy'))"
tail -3 gpurun_out/bench_${N}gpu.err
