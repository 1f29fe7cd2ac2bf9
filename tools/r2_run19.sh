#!/bin/bash
# state check after the schedule rebalance: GPU tests, smoke, bench line (default flags), e2e chunk 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r19_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r19_gputests.log
tail -3 gpurun_out/r19_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r19_smoke.log 2>&1; tail -2 gpurun_out/r19_smoke.log
timeout 900 python bench.py > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/r19_bench.json; echo; tail -3 gpurun_out/r19_bench.err
VRGDG_BENCH_CHUNK=1 timeout 900 python bench.py --no-cpu --no-extra --steps 5 > gpurun_out/r19_bench_chunk1.json 2> gpurun_out/r19_bench_chunk1.err; echo "bench chunk1 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r19_bench.json", "gpurun_out/r19_bench_chunk1.json"):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], json.dumps(d["e2e"])[:420])
PY
