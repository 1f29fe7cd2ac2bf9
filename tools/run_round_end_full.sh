bash tools/run_round_end_check.sh
timeout 400 python tools/quick_perf.py > gpurun_out/final_perf.log 2>&1; grep -E "unsharp_tma/1080p_f16|chain_g_u/1080p_f16|chain_g_l_u/1080p_f16/nat|enhancer" gpurun_out/final_perf.log
