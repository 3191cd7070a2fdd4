#!/bin/bash
# quick correctness + microbench (no profiler)
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
echo "== microbench"; timeout 600 python scripts/microbench.py "$@" 2>&1 | tail -8 | tee gpurun_out/microbench.log
