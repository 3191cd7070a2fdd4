#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
python scripts/microbench.py dec chain
python scripts/trace_probe.py dec 2>/dev/null | head -5
for cfg in "4 11" "6 11" "8 11" "10 11" "12 11" "12 10" "6 5" "4 5" "5 5" "6 8" "8 8"; do set -- $cfg
  echo "== warps=$1 stages=$2"; ORION_B200_WARPS=$1 ORION_B200_STAGES=$2 python scripts/microbench.py dec chainfm 2>&1 | grep " us "; done
