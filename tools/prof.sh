#!/bin/bash
# usage (on the GPU box): tools/prof.sh <tag> <kernel-regex> <skip> <kbench case...>
# Captures one `ncu --set full` launch and leaves only compact CSV exports (raw page + per-SASS-instruction source page)
# in gpurun_out/ (the .ncu-rep files embed the whole cubin and blow the 64 MiB return budget).
tag=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
rep=/tmp/${tag}.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f -o /tmp/${tag} \
    python tools/kbench.py "$@" > gpurun_out/${tag}.log 2>&1
ncu -i $rep --page raw --csv > gpurun_out/${tag}.raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source sass > /tmp/${tag}.sass.csv 2>/dev/null
gzip -9 -c /tmp/${tag}.sass.csv > gpurun_out/${tag}.sass.csv.gz
ncu -i $rep --page details > gpurun_out/${tag}.details.txt 2>/dev/null
ls -la gpurun_out/${tag}.*
