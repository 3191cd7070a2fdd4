#!/bin/bash
# tests + bench + ncu launch list + one ncu --set full capture of the chain kernel (B200_PROFILING.md recipe)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 30 -c 2 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
