#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -4 gpurun_out/r2_gputests.log
timeout 400 python tools/r2_perf.py chains > gpurun_out/r2_perf_em.jsonl 2> gpurun_out/r2_perf_em.err; echo "perf rc=$?"; cut -c1-120 gpurun_out/r2_perf_em.jsonl | grep -E "chain_g_l_u"
bash tools/r2_ncu.sh em_glu_f32 k_tile_em glu f32 2160 3840 4
