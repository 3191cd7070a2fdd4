#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r02_pytest_s13.log
echo "== microbench all"; timeout 400 python scripts/microbench.py dec chain fm rot lp configs 2>&1 | tee gpurun_out/r02_microbench_s13.txt | cut -c1-100
echo "== c2 launch list"
WARM_S=0.3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_kernel|osc_expand" -c 24 --csv --log-file gpurun_out/r02_launches_c2_s13.csv python scripts/microbench.py c2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2_s13.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[-6:]: print("  ", r[ki][:50], r[gi], r[bi], r[vi], "ns")
PY
echo "== trace C1 single launch"; ORION_B200_LIB=$PWD/orion-sdr_b200/variants/liborion_b200_hot.so timeout 200 python scripts/trace_probe.py chain 2>&1 | tee gpurun_out/s13_trace_c1.txt | head -12
