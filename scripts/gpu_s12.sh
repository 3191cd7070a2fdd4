#!/bin/bash
for m in 1 0; do
echo "== OSC_STREAM=$m"; ORION_B200_OSC_STREAM=$m timeout 300 python scripts/c2_probe.py firssb c2 2>&1 | tail -2
ORION_B200_OSC_STREAM=$m timeout 200 python scripts/microbench.py rot 2>&1 | tail -1 | cut -c1-90
done
