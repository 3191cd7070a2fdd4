#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload c5 --channels 128 --steps 2 --warmup 2 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_fir_kernel -s 2 -c 1 -f -o gpurun_out/r02_bank128 $CMD > gpurun_out/ncu_bank128.log 2>&1
echo "ncu exit=$?"
bash scripts/ncu_export.sh r02_bank128
ls -la gpurun_out | grep bank128
