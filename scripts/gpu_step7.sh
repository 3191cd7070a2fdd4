#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 700 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/s7_pytest.log
echo "== configs"; timeout 300 python scripts/microbench.py chain c3 c4 2>&1 | tail -4; for sp in 1 2; do echo "split $sp"; ORION_B200_SPLIT=$sp timeout 100 python scripts/microbench.py c4 2>&1 | tail -1; done
echo "== e2e"; timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c5 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(l['e2e']['value'], l['e2e']['pageable'])"
