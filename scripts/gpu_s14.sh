#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "overlapped_calls_near" 2>&1 | grep -E "passed|failed|Error|assert|watchdog" | head -8
done 2>&1 | tee gpurun_out/s14_flaky.txt
