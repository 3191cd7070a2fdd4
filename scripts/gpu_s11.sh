#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== c2 parts"; timeout 300 python scripts/c2_probe.py fir firssb0 firssb c2closed c2 2>&1 | tail -6 | tee gpurun_out/s11_c2_parts.txt
echo "== c2 big tiles"; ORION_B200_BIG_TILES=1 timeout 300 python scripts/c2_probe.py c2closed c2 2>&1 | tail -2
echo "== c2 no overlap"; OVERLAP=0 timeout 300 python scripts/microbench.py c2 2>&1 | tail -1
echo "== rot c2 (microbench)"; timeout 200 python scripts/microbench.py rot c2 2>&1 | tail -2
