#!/bin/bash
# att2all2 variant: kernels, golden fixture, oracle parity, gradients
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -q -m gpu -k "att2all2 or lstm_maxout or golden" 2>&1 | tail -30
