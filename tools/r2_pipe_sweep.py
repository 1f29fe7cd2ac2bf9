"""Pipelined colour-match schedule sweep (GPU box only): headline chain on 64 x 4K fp32 frames for each library under lib/variants
(+ stock), group sizes and tile-CTA limits.   python tools/r2_pipe_sweep.py"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "comfyui-vrgamedevgirl_b200")
CHILD = r'''
import importlib, os, sys, json, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, os.path.join(%(root)r, "tools"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import LUTS, natural_frames
from _clocks import Clocks
nv, ops = pkg._native, pkg.ops
dev = torch.device("cuda", 0)
CLK = Clocks(0)
lut = pkg.VRGDG_LUTS._parse_cube_file(os.path.join(LUTS, "B200 Vintage 33.cube"))
B, H, W = 64, 2160, 3840
x = natural_frames(8, H, W, seed=1, device=dev).repeat(B // 8, 1, 1, 1).contiguous()
out = torch.empty_like(x)
ref_sums = ops.lab_moments(natural_frames(1, H, W, seed=9, device=dev))
c = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), colormatch=dict(ref_sums=ref_sums, strength=1.0),
                        lut=dict(lut_data=lut, strength=10.0), stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=dev)
def timeit(fn, iters=6, warm=2):
    for _ in range(warm): fn()
    def run():
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]
    return CLK.sample_while(run)
for (g, ctas, serial) in %(cases)s:
    c.group_frames, c.serial = g, serial
    if ctas: os.environ["VRGDG_PIPE_TILE_CTAS"] = str(ctas)
    else: os.environ.pop("VRGDG_PIPE_TILE_CTAS", None)
    ms, clocks = timeit(lambda: c(x, out=out))
    print(json.dumps({"lib": os.path.basename(os.environ.get("VRGDG_B200_LIB", "stock")), "group_frames": g, "tile_ctas": ctas, "serial": serial, "frames": B,
                      "ms": round(ms, 3), "GPx/s": round(B * H * W / ms / 1e6, 2), "checksum": float(out[::8, ::64, ::64].sum().item()), "clocks": clocks}), flush=True)
'''


def main():
    libs = [None] + sorted(glob.glob(os.path.join(PKG, "lib", "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        cases = [(8, 0, False)]
        if lib:
            env["VRGDG_B200_LIB"] = lib
            cases = [(8, 0, False), (8, 0, True), (4, 0, False), (16, 0, False), (8, 280, False), (8, 264, False)]
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "cases": repr(cases)}], env=env, capture_output=True, text=True, timeout=900)
        print(r.stdout.strip() or ("FAILED %s: %s" % (lib, r.stderr[-600:])), flush=True)


if __name__ == "__main__":
    main()
