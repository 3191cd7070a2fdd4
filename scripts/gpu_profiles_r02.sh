#!/bin/bash
# Round-2 evidence capture: per-tile phase trace of the C1 chain, then `ncu --set full` of one steady-state launch of
# every kernel DESIGN.md quotes (C1 chain, channel-bank front end, C4 long-tap decimator, C3 chain, rate-1 FM, LpCascade).
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
    local name=$1 rx=$2 skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "$name: plain run failed"; tail -3 gpurun_out/plain_$name.log; return; }
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/r02_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "$name: ncu exit=$?"
}
if [ -f orion-sdr_b200/variants/liborion_b200_hot.so ]; then
  echo "== trace probe (C1 chain)"; ORION_B200_LIB=$PWD/orion-sdr_b200/variants/liborion_b200_hot.so timeout 300 python scripts/trace_probe.py chain 2>&1 | tee gpurun_out/r02_trace_chain.txt | tail -40
fi
cap c1chain chain_kernel 30 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
cap bank bank_fir_kernel 2 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu-baseline
WARM_S=0.3 cap c4 chain_kernel 12 python scripts/microbench.py c4
WARM_S=0.3 cap c3 chain_kernel 20 python scripts/microbench.py c3
WARM_S=0.3 cap fm chain_kernel 20 python scripts/microbench.py fm
WARM_S=0.3 cap lp chain_kernel 20 python scripts/microbench.py lp
ls -la gpurun_out/*.ncu-rep
