#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/att_trace.py 2>&1 | tail -40
