#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
echo "== microbench"; timeout 600 python scripts/microbench.py 2>&1 | tail -8
python scripts/microbench.py dec > gpurun_out/plain_dec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 1 -o gpurun_out/prof_dec python scripts/microbench.py dec > gpurun_out/ncu_dec.log 2>&1
echo "ncu dec exit=$?"
python scripts/microbench.py chainfm > gpurun_out/plain_chain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 1 -o gpurun_out/prof_chain python scripts/microbench.py chainfm > gpurun_out/ncu_chain.log 2>&1
echo "ncu chain exit=$?"
