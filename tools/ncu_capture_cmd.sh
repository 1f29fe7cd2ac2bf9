#!/bin/bash
# usage: tools/ncu_capture_cmd.sh <tag> <kernel-regex> <skip> <cmd...>   (GPU box; one launch, CSV exports into gpurun_out/)
tag=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip -c > gpurun_out/ncu_${tag}_source.csv.gz
ls -la /tmp/prof_$tag.ncu-rep gpurun_out/ncu_${tag}_* >> gpurun_out/ncu_$tag.log
