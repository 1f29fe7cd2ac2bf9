#!/bin/bash
# final evidence run of the round (1 GPU): parity, smoke, both bench arms, launch list, --set full captures, micro-benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/f_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_gputests.log
tail -3 gpurun_out/f_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/f_bench_reference.json
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/f_bench.json; echo; tail -3 gpurun_out/f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 600 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/f_launches_bench.log 2>&1; echo "launch list rc=$?"
bash tools/r2_ncu.sh apply_f32 k_tile full f32 2160 3840 8
bash tools/r2_ncu.sh momstore_f32 k_lab_moments full f32 2160 3840 8
bash tools/r2_ncu.sh configs1_f16 k_tile glu f16 1080 1920 64
timeout 900 python tools/r2_perf.py cm chains luts ext > gpurun_out/f_perf.jsonl 2> gpurun_out/f_perf.err; echo "perf rc=$?"; wc -l gpurun_out/f_perf.jsonl
ls gpurun_out | wc -l
