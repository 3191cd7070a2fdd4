#!/bin/bash
# Round capture: GPU tests, bench, ncu launch list, ncu --set full of the chain kernel (bench command) and of the
# decimator-only kernel (microbench).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-2500
echo "== microbench"; timeout 600 python scripts/microbench.py 2>&1 | tail -8 | tee gpurun_out/microbench.log
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 30 -c 1 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
python scripts/microbench.py dec > gpurun_out/plain_dec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 40 -c 1 -o gpurun_out/prof_dec python scripts/microbench.py dec > gpurun_out/ncu_dec.log 2>&1
echo "ncu dec exit=$?"
