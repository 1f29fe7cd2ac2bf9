#!/bin/bash
# round 2 final evidence run: parity, smoke, bench line, launch list, --set full captures, L2 evidence, micro-benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -4 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r2_launches_bench.log 2>&1; echo "launch list rc=$?"
bash tools/r2_ncu.sh apply_f32 k_tile full f32 2160 3840 8
bash tools/r2_ncu.sh momstore_f32 k_lab_moments full f32 2160 3840 8
bash tools/r2_ncu.sh configs1_f16 k_tile glu f16 1080 1920 64
VRGDG_G=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2_cm_1080p_G1.csv python tools/r2_prof_target.py cmg1 f32 1080 1920 4 > /dev/null 2>&1
VRGDG_G=4 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2_cm_1080p_G4.csv python tools/r2_prof_target.py cmg1 f32 1080 1920 4 > /dev/null 2>&1
timeout 900 python tools/r2_perf.py cm chains luts ext > gpurun_out/r2_perf.jsonl 2> gpurun_out/r2_perf.err; echo "perf rc=$?"; wc -l gpurun_out/r2_perf.jsonl
ls gpurun_out | wc -l
