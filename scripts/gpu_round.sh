#!/bin/bash
# Round capture: smoke, GPU tests, bench line, block / config microbenchmarks, ncu launch list of the bench command and one
# `ncu --set full` launch of the C1 chain kernel (exported to CSV on the box).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -x > gpurun_out/r02_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log > gpurun_out/r02_bench_line.json; cut -c1-400 gpurun_out/r02_bench_line.json
echo "== microbench"; timeout 900 python scripts/microbench.py dec chain fm rot lp configs 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r02_microbench.txt
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_kernel|bank_fir|osc_expand|agc_|fm_|ssb_|slice|gain" -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 30 -c 1 -f -o gpurun_out/r02_c1chain $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; bash scripts/ncu_export.sh r02_c1chain
ls -la gpurun_out | tail -12
