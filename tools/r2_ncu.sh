#!/bin/bash
# usage: tools/r2_ncu.sh <tag> <kernel-regex> <r2_prof_target args...>   (GPU box; writes small CSVs into gpurun_out/)
tag=$1; regex=$2; shift 2
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o /tmp/prof_$tag python tools/r2_prof_target.py "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > /tmp/ncu_${tag}_source.csv 2>/dev/null
gzip -c /tmp/ncu_${tag}_source.csv > gpurun_out/ncu_${tag}_source.csv.gz
ls -la /tmp/prof_$tag.ncu-rep gpurun_out/ncu_${tag}_* >> gpurun_out/ncu_$tag.log
