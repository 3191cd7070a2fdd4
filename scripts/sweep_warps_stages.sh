#!/bin/bash
for cfg in "4 4" "8 5" "8 10" "12 10" "16 6" "16 10" "16 11"; do set -- $cfg
  echo "== warps=$1 stages=$2"
  WARM_S=0.5 ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "
done
