#!/bin/bash
V=$PWD/orion-sdr_b200/variants
for i in 1 2; do
echo "== default";  timeout 100 python scripts/microbench.py chainfm 2>&1 | tail -1
for v in q2 q4 tapsmem; do echo "== $v"; ORION_B200_LIB=$V/liborion_b200_$v.so timeout 100 python scripts/microbench.py chainfm 2>&1 | tail -1; done
done
