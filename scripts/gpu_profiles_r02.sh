#!/bin/bash
# Round-2 evidence capture: `ncu --set full` of one steady-state launch of every kernel DESIGN.md quotes (C1 chain,
# channel-bank front end, C4 long-tap decimator, C3 chain, rate-1 FM, LpCascade), exported to CSV on the box
# (gpurun_out/ only travels back up to 64 MiB, so the .ncu-rep files are dropped after the export).
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    timeout 200 "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_$name.log; return; }
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name: ncu exit=$?"
    bash scripts/ncu_export.sh r02_$name
}
for w in "$@"; do case $w in
  c1)   cap c1chain chain_kernel 30 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 ;;
  bank) cap bank bank_fir_kernel 2 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu-baseline ;;
  c4)   WARM_S=0.3 cap c4 chain_kernel 12 python scripts/microbench.py c4 ;;
  c2)   WARM_S=0.3 cap c2 chain_kernel 8 python scripts/microbench.py c2 ;;
  c3)   WARM_S=0.3 cap c3 chain_kernel 20 python scripts/microbench.py c3 ;;
  fm)   WARM_S=0.3 cap fm chain_kernel 20 python scripts/microbench.py fm ;;
  lp)   WARM_S=0.3 cap lp chain_kernel 20 python scripts/microbench.py lp ;;
esac; done
ls -la gpurun_out | tail -30
