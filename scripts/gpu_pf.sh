#!/bin/bash
# L2 look-ahead of the stage ring: sweep of the prefetch distance on the unified and the warp-specialised C1 kernels
echo "== quick check"; timeout 100 python -m pytest tests/test_golden.py -m gpu -q --timeout 60 -p no:cacheprovider -x 2>&1 | tail -2
for pf in 0 8 16 24 40 64; do
  echo "== unified pf=$pf"; ORION_B200_L2_PREFETCH=$pf ORION_B200_NO_WS=1 timeout 60 python scripts/microbench.py dec chainfm 2>&1 | tail -2
done
for pf in 0 16 24 40; do
  echo "== ws pf=$pf"; ORION_B200_L2_PREFETCH=$pf timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
done
echo "== unified pf=24 isolated"; OVERLAP=0 ORION_B200_L2_PREFETCH=24 ORION_B200_NO_WS=1 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== c3/c4 pf default"; ORION_B200_NO_WS=1 timeout 100 python scripts/microbench.py c3 c4 2>&1 | tail -2
echo "== c3/c4 pf 0"; ORION_B200_L2_PREFETCH=0 ORION_B200_NO_WS=1 timeout 100 python scripts/microbench.py c3 c4 2>&1 | tail -2
