#!/bin/bash
# usage (on the GPU box): tools/prof2.sh <tag> <kernel-regex> <skip> <nwarps> <kbench case...>
# One `ncu --set full` launch of the named kernel; leaves the text summary (tools/ncu_summary.py) and the gzipped
# per-SASS-instruction source page in gpurun_out/ (the .ncu-rep itself embeds the cubin and is too big to return).
tag=$1; regex=$2; skip=$3; nw=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f -o /tmp/${tag} \
    python tools/kbench.py "$@" > gpurun_out/${tag}.log 2>&1
python tools/ncu_summary.py /tmp/${tag}.ncu-rep ${nw} > gpurun_out/${tag}.txt 2>&1
ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/${tag}.sass.csv.gz
cat gpurun_out/${tag}.txt
