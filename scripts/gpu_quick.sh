#!/bin/bash
# quick correctness + bench (no profiler)
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-3000
