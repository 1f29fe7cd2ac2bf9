#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r14_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r14_gputests.log
tail -4 gpurun_out/r14_gputests.log
python tools/r2_variants.py all 2>&1 | tee gpurun_out/r14_variants.jsonl | cut -c1-120
