#!/bin/bash
# what the driver runs at round end: the whole GPU suite in ONE process, smoke(), the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench exit $?"; wc -l gpurun_out/bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','gpu_launches','dtype','vs_baseline')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks'])"
