"""Times the headline fused chain (grain -> 33^3 LUT -> unsharp, fp16 1080p) for every library under lib/variants plus the stock
one (GPU box only; tuning experiments).  python tools/variant_perf.py [frames] [chain|unsharp_f16|unsharp_f32]"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "comfyui-vrgamedevgirl_b200")

CHILD = r'''
import importlib, os, sys, json, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import LUTS, natural_frames
nv = pkg._native
dev = torch.device("cuda", 0)
frames = %(frames)d
x = natural_frames(frames, 1080, 1920, seed=1, dtype=torch.float16, device=dev)
out = torch.empty_like(x)
lut = pkg.VRGDG_LUTS._parse_cube_file(os.path.join(LUTS, "B200 Vintage 33.cube"))
mode = %(mode)r
if mode.startswith("unsharp"):
    if mode.endswith("f32"):
        x = x.float(); out = torch.empty_like(x)
    chain = lambda x, out=None: pkg.ops.stencil3x3(x, nv.STENCIL_BOX_UNSHARP, 0.5, nv.BORDER_REPLICATE)
else:
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), lut=dict(lut_data=lut, strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): chain(x, out=out)
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); chain(x, out=out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
print(json.dumps({"lib": os.path.basename(os.environ.get("VRGDG_B200_LIB", "stock")), "mode": mode, "frames": frames, "ms": round(ms, 4),
                  "GPx/s": round(frames * 1080 * 1920 / ms / 1e6, 2), "checksum": float(out.float().sum().item())}))
'''


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    mode = sys.argv[2] if len(sys.argv) > 2 else "chain"
    libs = [None] + sorted(glob.glob(os.path.join(PKG, "lib", "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["VRGDG_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "frames": frames, "mode": mode}], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or ("FAILED %s: %s" % (lib, r.stderr[-400:])), flush=True)


if __name__ == "__main__":
    main()
