#!/bin/bash
# ncu_export.sh NAME -- on the GPU box: turn gpurun_out/NAME.ncu-rep into the two CSV pages the summary scripts read
# (raw counters, per-source-line samples), gzip them and drop the report itself (gpurun_out/ travels back only up to 64 MiB).
n=$1
ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null
ncu -i gpurun_out/$n.ncu-rep --page source --csv > gpurun_out/$n.sass.csv 2>/dev/null
ncu -i gpurun_out/$n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/$n.cuda.csv 2>/dev/null
gzip -f gpurun_out/$n.raw.csv gpurun_out/$n.sass.csv gpurun_out/$n.cuda.csv
rm -f gpurun_out/$n.ncu-rep
