#!/bin/bash
# 2 GPUs: the NCCL test with world_size 2, the bench under torchrun (all-gather inside every step), the reference arm
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r2_nccl2.log 2>&1; echo "nccl test rc=$?"; tail -3 gpurun_out/r2_nccl2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/r2_bench_2gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_reference.json
