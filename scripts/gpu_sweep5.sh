#!/bin/bash
echo "== tests"; timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -3
echo "== idle stage on";  timeout 200 python scripts/microbench.py chain c3 2>&1 | tail -3
echo "== idle stage off"; ORION_B200_NO_IDLE_STAGE=1 timeout 200 python scripts/microbench.py chain c3 2>&1 | tail -3
