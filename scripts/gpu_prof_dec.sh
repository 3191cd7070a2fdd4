#!/bin/bash
mkdir -p gpurun_out
export ORION_B200_WARPS=2 ORION_B200_STAGES=5
python scripts/microbench.py dec > gpurun_out/plain_dec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 1 -o gpurun_out/prof_dec3 python scripts/microbench.py dec > gpurun_out/ncu_dec.log 2>&1
echo "ncu dec exit=$?"; cat gpurun_out/plain_dec.log | tail -2
