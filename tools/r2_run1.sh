#!/bin/bash
# round 2, GPU call 1: parity (all -m gpu tests), the re-pointed bench line, the micro-benchmark rows of this round
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -5 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
timeout 500 python tools/r2_perf.py cm chains luts > gpurun_out/r2_perf.jsonl 2> gpurun_out/r2_perf.err; echo "perf rc=$?"; cut -c1-160 gpurun_out/r2_perf.jsonl
