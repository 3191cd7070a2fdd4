#!/bin/bash
for v in 111 000 100 001 110; do
  for cfg in "8 5" "16 10"; do set -- $cfg
    echo "== variant $v (packed,taps_smem,hot_smem) warps=$1 stages=$2"
    ORION_B200_LIB=$PWD/orion-sdr_b200/variants/liborion_b200_$v.so WARM_S=0.6 ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "
  done
done
