#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
python scripts/trace_probe.py dec 2>&1 | sed -n 3,6p
python scripts/trace_probe.py chain 2>&1 | sed -n 1,12p
for cfg in "8 11" "10 11" "12 11" "12 10" "12 9" "12 8" "6 5" "5 5"; do set -- $cfg
  echo "== warps=$1 stages=$2"; WARM_S=0.7 ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "; done
