#!/bin/bash
for cfg in "2 5" "3 5" "4 5" "5 5" "4 8" "6 8" "8 8" "3 4" "2 4" "2 3"; do set -- $cfg
  echo "== warps=$1 stages=$2"; ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "; done
