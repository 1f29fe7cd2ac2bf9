#!/bin/bash
# round 2, GPU call: parity with the element-mapped tile kernel, then A/B of the chains (new vs VRGDG_TILE_LEGACY=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -15 gpurun_out/r2_gputests.log
timeout 400 python tools/r2_perf.py cm chains > gpurun_out/r2_perf_em.jsonl 2> gpurun_out/r2_perf_em.err; echo "perf rc=$?"; cut -c1-120 gpurun_out/r2_perf_em.jsonl | grep -E "chain|full|moments|colormatch"
VRGDG_TILE_LEGACY=1 timeout 400 python tools/r2_perf.py cm chains > gpurun_out/r2_perf_legacy.jsonl 2> gpurun_out/r2_perf_legacy.err; echo "perf rc=$?"; cut -c1-120 gpurun_out/r2_perf_legacy.jsonl | grep -E "chain_g_l_u|full"
