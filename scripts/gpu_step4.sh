#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 600 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/s4_pytest.log
echo "== configs"; timeout 200 python scripts/microbench.py chain c3 c4 2>&1 | tail -4
echo "== c5 small banks"; for ch in 128 1024; do timeout 120 python bench.py --workload c5 --channels $ch --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['config']['channels_per_gpu'], 'channels', round(l['ms_per_step'],3), 'ms', l['parity_check'])"; done
echo "== e2e threads"; for t in 4 8 12; do ORION_B200_COPY_THREADS=$t timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c5 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('threads $t', l['e2e']['value'], l['e2e']['pageable'])"; done
nproc
