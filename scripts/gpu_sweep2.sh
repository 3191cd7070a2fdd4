#!/bin/bash
for st in 10 11 9; do echo "== stages $st"; ORION_B200_STAGES=$st timeout 100 python scripts/microbench.py dec chainfm 2>&1 | tail -2; done
echo "== stages 11 again"; ORION_B200_STAGES=11 timeout 100 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== default again"; timeout 100 python scripts/microbench.py chainfm 2>&1 | tail -1
