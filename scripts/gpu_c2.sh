#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/microbench.py c2"
WARM_S=0.3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_kernel|osc_expand" -c 30 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[-9:]: print("  ", r[ki][:50], r[gi], r[bi], r[vi], "ns")
PY
bash scripts/gpu_profiles_r02.sh c2
