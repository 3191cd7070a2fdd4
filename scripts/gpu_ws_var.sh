#!/bin/bash
echo "== quick check"; timeout 100 python -m pytest tests/test_golden.py -m gpu -q --timeout 60 -p no:cacheprovider -x 2>&1 | tail -2
echo "== ws (default)";      timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== unified (ORION_B200_NO_WS)"; ORION_B200_NO_WS=1 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
V=$PWD/orion-sdr_b200/variants
for v in $(ls $V 2>/dev/null | sed 's/liborion_b200_//; s/\.so//'); do echo "== $v"; ORION_B200_LIB=$V/liborion_b200_$v.so timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1; done
echo "== ws (default) again"; timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== ws 96M"; N_SAMPLES=96000000 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
echo "== unified 96M"; ORION_B200_NO_WS=1 N_SAMPLES=96000000 timeout 60 python scripts/microbench.py chainfm 2>&1 | tail -1
