#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/s3_pytest.log
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/s3_bench.json | cut -c1-3500
echo "== rate-1 blocks, L2 prefetch off / 32 / 64"
for pf in 0 32 64; do ORION_B200_L2_PREFETCH=$pf timeout 100 python scripts/microbench.py fm lp 2>&1 | tail -2; done
echo "== c5 small banks"; for ch in 128 256 512 1024; do timeout 120 python bench.py --workload c5 --channels $ch --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['config']['channels_per_gpu'], 'channels', round(l['ms_per_step'],3), 'ms', l['parity_check'])"; done
