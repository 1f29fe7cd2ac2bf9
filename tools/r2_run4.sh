#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -12 gpurun_out/r2_gputests.log
timeout 600 python tools/r2_perf.py cm > gpurun_out/r2_perf_cm.jsonl 2> gpurun_out/r2_perf_cm.err; echo "perf rc=$?"; cut -c1-110 gpurun_out/r2_perf_cm.jsonl; tail -3 gpurun_out/r2_perf_cm.err
