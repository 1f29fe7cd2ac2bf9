#!/bin/bash
# round 2: parity, the bench line, its launch list, --set full captures of the two kernels of the headline step, L2 evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -4 gpurun_out/r2_gputests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r2_launches_bench.log 2>&1; echo "launch list rc=$?"
bash tools/r2_ncu.sh apply_f32 k_tile full f32 2160 3840 8
bash tools/r2_ncu.sh momstore_f32 k_lab_moments full f32 2160 3840 8
VRGDG_G=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2_cm_1080p_G1.csv python tools/r2_prof_target.py cmg1 f32 1080 1920 4 > /dev/null 2>&1
VRGDG_G=4 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r2_cm_1080p_G4.csv python tools/r2_prof_target.py cmg1 f32 1080 1920 4 > /dev/null 2>&1
ls -la gpurun_out | tail -20
