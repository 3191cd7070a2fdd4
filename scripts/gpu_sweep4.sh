#!/bin/bash
python scripts/microbench.py
for cfg in "6 11" "8 11" "10 11" "12 11" "12 10" "6 5" "8 5"; do set -- $cfg
  echo "== warps=$1 stages=$2"; ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "; done
