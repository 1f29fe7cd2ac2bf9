mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/pytest29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest29.log; tail -3 gpurun_out/pytest29.log
timeout 400 python tools/quick_perf.py > gpurun_out/perf29.log 2>&1; grep -E "unsharp_tma|chain_g_l_u/1080p_f16|enhancer" gpurun_out/perf29.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench29.log 2>&1; tail -1 gpurun_out/bench29.log | cut -c1-300
bash tools/ncu_capture_cmd.sh bench_chain k_tile 3 python bench.py --steps 2 --warmup 1 --no-cpu
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench29.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches29.log 2>&1
bash tools/ncu_capture.sh clarity9 k_adjust_box clarity f32 4
bash tools/ncu_capture.sh bicubic k_resize bicubic f32 4
bash tools/ncu_capture.sh lanczos_h k_lanczos_h lanczos f32 4
bash tools/ncu_capture.sh lanczos_v k_lanczos_v lanczos f32 4
bash tools/ncu_capture.sh unsharp_f32 k_tile unsharp f32 8
rm -f gpurun_out/*source.csv.gz.tmp; ls -la gpurun_out | tail -30; du -sh gpurun_out
