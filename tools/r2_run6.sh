#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -12 gpurun_out/r2_gputests.log
timeout 400 python tools/r2_perf.py chains > gpurun_out/r2_perf_coop.jsonl 2> gpurun_out/r2_perf_coop.err; echo "perf rc=$?"; cut -c1-120 gpurun_out/r2_perf_coop.jsonl | grep -E "chain_g_l_u"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_bench_coop.json 2> gpurun_out/r2_bench_coop.err; cut -c1-200 gpurun_out/r2_bench_coop.json
bash tools/r2_ncu.sh coop_glu_f32 k_tile glu f32 2160 3840 4
