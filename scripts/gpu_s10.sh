#!/bin/bash
mkdir -p gpurun_out
echo "== tests (mix / exact / ssb / rotator / c2)"; timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "rot or nco or ssb or usb or c2 or mix or exact or osc or modul" 2>&1 | tail -3
echo "== c2 parts"; timeout 300 python scripts/c2_probe.py fir firssb0 firssb c2closed c2 2>&1 | tail -6 | tee gpurun_out/s10_c2_parts.txt
echo "== c2 parts ROWR=1"; ORION_B200_ROWR=1 timeout 300 python scripts/c2_probe.py fir firssb0 c2closed c2 am25 2>&1 | tail -6 | tee gpurun_out/s10_c2_parts_r1.txt
echo "== rot"; timeout 200 python scripts/microbench.py rot 2>&1 | tail -1
