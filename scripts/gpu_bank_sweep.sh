#!/bin/bash
run() { timeout 120 python bench.py --workload c5 --channels $1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$*', round(l['ms_per_step'],3), 'ms', l['parity_check']['pass'])"; }
run 128
for cfg in "3 10" "4 12" "4 8" "6 8" "3 20" "2 24" "3 8"; do set -- $cfg; ORION_B200_BANK_NS=$1 ORION_B200_BANK_BT=$2 run 128 NS=$1 BT=$2; done
for r in 444 592 296; do ORION_B200_BANK_RANGES=$r run 128 ranges=$r; done
