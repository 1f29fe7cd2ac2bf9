#!/bin/bash
# N GPUs (default 8): the bench under torchrun exactly as the driver launches it (all-gather inside every step; e2e leg on every rank)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8; nvidia-smi topo -m 2>/dev/null | head -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${N}gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','clocks')}); print(json.dumps(d.get('e2e'))[:900])
PY
tail -3 gpurun_out/r2_bench_${N}gpu.err
