#!/bin/bash
# two back-to-back short bench runs (noise estimate)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ab$i.json')); print({k:round(d[k],1) for k in ('value','greedy_captions_per_s')}, d['train'] and (round(d['train']['value']), round(d['train']['ms_per_step'],3)))"
done
