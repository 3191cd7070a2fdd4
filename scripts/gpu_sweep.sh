#!/bin/bash
python scripts/microbench.py 2>&1 | tail -8
for w in 4 8 12 16; do for s in 2 3 4 5 6; do
  echo "== warps=$w stages=$s"; ORION_B200_WARPS=$w ORION_B200_STAGES=$s python scripts/microbench.py dec chain 2>&1 | grep -E "us " ; done; done
