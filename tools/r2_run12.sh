#!/bin/bash
# state check at the start of a session: GPU tests, then the default bench line (what the driver runs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r12_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r12_gputests.log
tail -4 gpurun_out/r12_gputests.log
timeout 900 python bench.py > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r12_bench.json; tail -3 gpurun_out/r12_bench.err
