#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -12 gpurun_out/r2_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_piped.json 2> gpurun_out/r2_bench_piped.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_piped.json; tail -3 gpurun_out/r2_bench_piped.err
timeout 900 python tools/r2_perf.py cm > gpurun_out/r2_perf_piped.jsonl 2> gpurun_out/r2_perf_piped.err; echo "perf rc=$?"; cut -c1-110 gpurun_out/r2_perf_piped.jsonl | grep -E "full_chain|one_call"; tail -3 gpurun_out/r2_perf_piped.err
