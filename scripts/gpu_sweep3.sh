#!/bin/bash
echo "== quick tests"; timeout 300 python -m pytest tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -q --timeout 100 -p no:cacheprovider -x 2>&1 | tail -2
V=$PWD/orion-sdr_b200/variants/liborion_b200_hot.so
for i in 1 2; do
echo "== early finish on";  timeout 100 python scripts/microbench.py chain lp 2>&1 | tail -3
echo "== early finish off"; ORION_B200_LIB=$V timeout 100 python scripts/microbench.py chain lp 2>&1 | tail -3
done
