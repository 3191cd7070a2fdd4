#!/bin/bash
mkdir -p gpurun_out
for ch in 128 1024; do
CMD="python bench.py --workload c5 --channels $ch --steps 2 --warmup 2 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bank_fir|chain_kernel|osc_expand" -c 40 --csv --log-file gpurun_out/r02_launches_c5_$ch.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c5_$ch.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
print("channels $ch")
for r in rows[-9:]: print("  ", r[ki][:50], r[gi], r[bi], r[vi], "ns")
PY
done
echo "== exact blocks, oscillator walked ahead"; timeout 300 python scripts/microbench.py rot c2 2>&1 | tail -2
