#!/bin/bash
mkdir -p gpurun_out
run() { timeout 200 python bench.py --workload c5 --channels $1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check'])"; }
echo "== bank tests"; timeout 300 python -m pytest tests/test_channel_bank.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider -k "bank or c5" 2>&1 | tail -2
echo "== c5 128 grid 1"; ORION_B200_BATCH_GRID=1 run 128
echo "== c5 128 default (2)"; run 128
echo "== c5 128 grid 4"; ORION_B200_BATCH_GRID=4 run 128
echo "== c5 256 default"; run 256
echo "== c5 256 grid1"; ORION_B200_BATCH_GRID=1 run 256
echo "== c5 1024"; run 1024
