#!/bin/bash
run() { timeout 200 python bench.py --workload c5 --channels $1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['max_err_fs'], d['parity_check']['pass'])"; }
echo "== bank tests"; timeout 300 python -m pytest tests/test_channel_bank.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider -k "bank or c5" 2>&1 | tail -2
echo "== c5 128 default"; run 128
echo "== c5 128 reverse"; ORION_B200_BANK_REVERSE=1 run 128
echo "== c5 128 BT=8"; ORION_B200_BANK_BT=8 run 128
echo "== c5 1024"; run 1024
