#!/bin/bash
mkdir -p gpurun_out
run() { timeout 200 python bench.py --workload c5 --channels 128 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['pass'])"; }
echo "== c5 128 default"; run
echo "== c5 128 no fork"; ORION_B200_BANK_NO_FORK=1 run
echo "== c5 128 BT=8 ranges 1184"; ORION_B200_BANK_BT=8 ORION_B200_BANK_RANGES=1184 run
echo "== c5 128 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bank_fir|chain_kernel" -c 18 --csv --log-file gpurun_out/r02_launches_c5_128.csv python bench.py --workload c5 --channels 128 --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c5_128.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[-6:]: print("  ", r[ki][:60], r[gi], r[bi], r[vi], "ns")
PY
