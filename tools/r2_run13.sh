#!/bin/bash
# compact LUT cells: GPU tests, bench (no CPU leg), kernel micro-benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r13_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r13_gputests.log
tail -15 gpurun_out/r13_gputests.log
timeout 900 python bench.py --no-cpu > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r13_bench.json; tail -3 gpurun_out/r13_bench.err
