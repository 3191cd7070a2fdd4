#!/bin/bash
# Final round-2 evidence from the final build: GPU tests, the default bench line, the ncu launch list of the bench command,
# `ncu --set full` of the C1 and C2 kernels, the microbenchmarks.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/r02_gpu_tests_final.log; cat gpurun_out/r02_gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_line_final.json 2> gpurun_out/r02_bench_final.err; tail -c 600 gpurun_out/r02_bench_line_final.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_line_reference_final.json 2>> gpurun_out/r02_bench_final.err; cat gpurun_out/r02_bench_line_reference_final.json | cut -c1-400
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/ncu_launches_final.log 2>&1
grep -c chain_kernel gpurun_out/r02_launches_bench_final.csv
bash scripts/gpu_profiles_r02.sh c1 2>&1 | grep "ncu exit"
timeout 400 python scripts/microbench.py dec chain fm rot lp configs 2>&1 | tee gpurun_out/r02_microbench_final.txt | cut -c1-100
