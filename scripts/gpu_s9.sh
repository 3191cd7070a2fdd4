#!/bin/bash
mkdir -p gpurun_out
echo "== bank tests"; timeout 300 python -m pytest tests/test_channel_bank.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
echo "== c2 parts"; timeout 300 python scripts/c2_probe.py fir firssb0 firssb c2closed c2 am25 fm25 2>&1 | tail -12 | tee gpurun_out/s9_c2_parts.txt
echo "== c5 128 default"; timeout 200 python bench.py --workload c5 --channels 128 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['pass'])"
for cfg in "8 592" "4 592" "16 592" "8 1184"; do set -- $cfg
echo "== c5 128 BT=$1 RANGES=$2"; ORION_B200_BANK_BT=$1 ORION_B200_BANK_RANGES=$2 timeout 200 python bench.py --workload c5 --channels 128 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity_check']['pass'])"
done
echo "== c5 128 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s9_launches_c5_128.csv python bench.py --workload c5 --channels 128 --steps 2 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/s9_launches_c5_128.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[-8:]: print("  ", r[ki][:60], r[gi], r[bi], r[vi], "ns")
PY
