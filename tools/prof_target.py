"""Runs one kernel family a few times so that ncu can capture it (GPU box only).
    python tools/prof_target.py {chain|lut|grain|unsharp|colormatch|clarity|bicubic|lanczos} [f16|f32] [frames] [nat|white]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import LUTS, natural_frames  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "chain"
dt = torch.float16 if (len(sys.argv) > 2 and sys.argv[2] == "f16") else torch.float32
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dist = sys.argv[4] if len(sys.argv) > 4 else "nat"
dev = torch.device("cuda", 0)
nv, ops = pkg._native, pkg.ops
x = natural_frames(frames, 1080, 1920, seed=1, dtype=dt, device=dev) if dist == "nat" else torch.rand(frames, 1080, 1920, 3, device=dev).to(dt)
lut = pkg.VRGDG_LUTS._parse_cube_file(os.path.join(LUTS, "B200 Vintage 33.cube"))
lut_dev = ops.pack_lut(lut["lut"], dev)
out = torch.empty_like(x)
if what == "chain":
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), lut=dict(lut_data=lut, strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=dev)
    fn = lambda: chain(x, out=out)
elif what == "lut":
    fn = lambda: ops.lut3d_apply(x, lut_dev, [0, 0, 0], [1, 1, 1], 1.0, 0.0)
elif what == "grain":
    fn = lambda: ops.grain(x, 0.04, 0.5, 0.5, seed=42)
elif what == "unsharp":
    fn = lambda: ops.stencil3x3(x, nv.STENCIL_BOX_UNSHARP, 0.5, nv.BORDER_REPLICATE)
elif what == "clarity":
    vt = importlib.import_module("comfyui-vrgamedevgirl_b200.video_tools")
    desc = vt._adjust_desc({"contrast": 10, "clarity": 50}, 1080, 1920)
    fn = lambda: ops.adjust(x, desc)
elif what == "bicubic":
    fn = lambda: ops.resize(x[:4], 2160, 3840, "bicubic")
elif what == "lanczos":
    u = (x[:4].float() * 255).to(torch.uint8)
    fn = lambda: ops.resize_lanczos4_u8(u, 2160, 3840)
else:
    sums = ops.lab_moments(x)
    params = ops.colormatch_params(sums, sums[:1].contiguous())
    fn = lambda: ops.colormatch_apply(x, params, 1.0, 0.0)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done", what, dt, frames, dist)
