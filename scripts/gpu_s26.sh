#!/bin/bash
run() { timeout 200 python bench.py --workload c5 --channels $1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['max_err_fs'], d['parity_check']['pass'])"; }
for i in 1 2; do
echo "== A (packed step, scalar mixer, 54 instr) 1024"; run 1024
echo "== C (scalar step, scalar mixer, 58 instr) 1024"; ORION_B200_LIB=$PWD/orion-sdr_b200/variants/liborion_b200_C.so run 1024
done
