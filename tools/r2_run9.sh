#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log
tail -4 gpurun_out/r2_gputests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
for g in 1; do
VRGDG_G=$g timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none --cache-control none -k regex:k_ -c 80 --csv --log-file gpurun_out/r2_cm_720p_G$g.csv python tools/r2_prof_target.py cmg1 f32 720 1280 6 > /dev/null 2>&1
done
