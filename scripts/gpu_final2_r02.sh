#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/r02_gpu_tests_final.log; cat gpurun_out/r02_gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_line_final.json 2> gpurun_out/r02_bench_final.err; tail -c 700 gpurun_out/r02_bench_line_final.json; echo
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_profiles_r02.sh bank 2>&1 | grep "ncu exit"
