#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload c5 --channels 128 --steps 2 --warmup 2 --no-cpu-baseline"
timeout 120 $CMD > gpurun_out/plain_c5_128.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c5_128.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c5_128.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); 
for r in rows[1:40]: print(r[ki][:60], r[vi])
PY
bash scripts/gpu_profiles_r02.sh c3
