#!/bin/bash
run() { timeout 200 python bench.py --workload c5 --channels $1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['max_err_fs'], d['parity_check']['pass'])"; }
echo "== bank tests"; timeout 300 python -m pytest tests/test_channel_bank.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider -k "bank or c5" 2>&1 | tail -2
echo "== c5 128 default (BT 12)"; run 128
echo "== c5 128 BT=16"; ORION_B200_BANK_BT=16 run 128
echo "== c5 128 BT=8 NS=4"; ORION_B200_BANK_BT=8 ORION_B200_BANK_NS=4 run 128
echo "== c5 1024"; run 1024
echo "== c5 1024 NS=3"; ORION_B200_BANK_NS=3 run 1024
echo "== c5 256"; run 256
