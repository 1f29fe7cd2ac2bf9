#!/bin/bash
mkdir -p gpurun_out
for c in 0 280 264 248 232 216 200; do
  echo "== tile CTAs $c"
  VRGDG_PIPE_TILE_CTAS=$c timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'])"
done
