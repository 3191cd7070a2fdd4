#!/bin/bash
for cfg in "5 3 296" "5 3 444" "4 2 444" "4 2 592" "4 2 296" "8 5 148" "8 5 296" "3 2 592" "3 2 740"; do set -- $cfg
  echo "== warps=$1 stages=$2 grid=$3 (overlap on)"
  WARM_S=0.5 ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 ORION_B200_GRID=$3 timeout 120 python scripts/microbench.py dec chainfm 2>&1 | grep " us "
done
