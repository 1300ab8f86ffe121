#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q --no-header -p no:cacheprovider --timeout 300 "$@" 2>&1 | tail -25
