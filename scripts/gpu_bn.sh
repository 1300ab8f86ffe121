#!/bin/bash
# use_bn = 1: statistics kernel + folded BatchNorm in att_embed
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -k "use_bn or col_moments" 2>&1 | tail -40
