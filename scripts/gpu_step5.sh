#!/bin/bash
mkdir -p gpurun_out
echo "== new tests first"; timeout 300 python -m pytest tests/test_gpu_next_rows.py tests/test_gpu_agc.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | tail -15
echo "== full gpu suite"; timeout 700 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -6 | tee gpurun_out/s5_pytest.log
echo "== configs"; timeout 200 python scripts/microbench.py chain c3 c4 2>&1 | tail -4
echo "== c5 small banks"; for ch in 128 1024; do timeout 120 python bench.py --workload c5 --channels $ch --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['config']['channels_per_gpu'], 'channels', round(l['ms_per_step'],3), 'ms', l['parity_check'])"; done
echo "== e2e"; timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c5 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['e2e']['value'], l['e2e']['pageable'])"
