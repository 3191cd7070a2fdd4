#!/bin/bash
# time of one launch against the stream length: the fixed per-launch cost and the steady-state rate
for o in 0 1; do for n in 6000000 12000000 24000000 48000000 96000000; do
  echo "== overlap=$o n=$n"; OVERLAP=$o N_SAMPLES=$n WARM_S=0.5 timeout 200 python scripts/microbench.py dec chainfm 2>&1 | grep " us "
done; done
