#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -15 gpurun_out/r2_gputests.log
timeout 400 python tools/r2_perf.py ext > gpurun_out/r2_perf_ext.jsonl 2> gpurun_out/r2_perf_ext.err; echo "perf rc=$?"; cut -c1-120 gpurun_out/r2_perf_ext.jsonl; tail -3 gpurun_out/r2_perf_ext.err
