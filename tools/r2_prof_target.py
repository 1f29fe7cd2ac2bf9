"""ncu target of round 2 (GPU box only): runs one workload a few times.
    python tools/r2_prof_target.py {full|glu|cmu|cmg1|moments|cmnode} {f32|f16} H W frames"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import LUTS, natural_frames  # noqa: E402

what = sys.argv[1]
dt = torch.float16 if sys.argv[2] == "f16" else torch.float32
H, W, frames = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
dev = torch.device("cuda", 0)
nv, ops = pkg._native, pkg.ops
x = natural_frames(frames, H, W, seed=1, dtype=dt, device=dev)
out = torch.empty_like(x)
lut = pkg.VRGDG_LUTS._parse_cube_file(os.path.join(LUTS, "B200 Vintage 33.cube"))
ref_sums = ops.lab_moments(natural_frames(1, H, W, seed=9, dtype=dt, device=dev))
G = dict(intensity=0.04, saturation_mix=0.5, seed=42)
S = dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5)
if what == "full":
    c = pkg.chain.PostChain(grain=G, colormatch=dict(ref_sums=ref_sums, strength=1.0), lut=dict(lut_data=lut, strength=10.0), stencil=S, device=dev)
    fn = lambda: c(x, out=out)
elif what == "glu":
    c = pkg.chain.PostChain(grain=G, lut=dict(lut_data=lut, strength=10.0), stencil=S, device=dev)
    fn = lambda: c(x, out=out)
elif what == "cmu":
    c = pkg.chain.PostChain(colormatch=dict(ref_sums=ref_sums, strength=1.0), stencil=S, device=dev)
    fn = lambda: c(x, out=out)
elif what == "cmg1":       # colour match alone, one frame per group: the f-planes of a group stay in L2 between the two passes
    c = pkg.chain.PostChain(colormatch=dict(ref_sums=ref_sums, strength=1.0), device=dev)
    c.group_frames = int(os.environ.get("VRGDG_G", "1"))
    fn = lambda: c(x, out=out)
elif what == "moments":
    d = nv.ChainDesc()
    d.grain_enabled, d.grain_intensity, d.grain_sat, d.grain_one_minus_sat, d.grain_seed = 1, 0.04, 0.5, 0.5, 42
    fn = lambda: ops.chain_lab_moments(x, d)
else:
    fn = lambda: ops.colormatch_apply(x, ops.colormatch_params(ops.lab_moments(x), ref_sums), 1.0, 0.0)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done", what, dt, H, W, frames)
