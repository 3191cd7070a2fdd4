#!/bin/bash
for cfg in "8 5" "8 4" "6 3" "5 3" "4 2" "16 10" "12 5"; do set -- $cfg
  echo "== warps=$1 stages=$2 (overlap on)"
  WARM_S=0.5 ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 timeout 120 python scripts/microbench.py dec chainfm 2>&1 | grep " us "
done
