"""Times the LUT chains for the stock library and every library under lib/variants (GPU box only; tuning experiments, see
tools/build_variant.sh).   python tools/r2_variants.py [f32|f16|all]
rows: glu = fused grain -> 33^3 LUT -> unsharp (one k_tile launch), full = the headline chain (grain -> colour match -> LUT -> unsharp)"""
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
G = dict(intensity=0.04, saturation_mix=0.5, seed=42)
S = dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=8, warm=3):
    for _ in range(warm): fn()
    def run():
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]
    return CLK.sample_while(run)
rows = []
what = %(what)r
cases = []
if what in ("f32", "all"): cases += [("glu", 8, 2160, 3840, torch.float32, "4k_f32"), ("full", 32, 2160, 3840, torch.float32, "4k_f32")]
if what == "locality":      # same instruction stream, different table locality: is the gather bound by L1 misses?
    cases += [("glu", 8, 2160, 3840, torch.float32, "4k_f32"), ("glu_lowgrain", 8, 2160, 3840, torch.float32, "4k_f32"), ("glu_flat", 8, 2160, 3840, torch.float32, "4k_f32"),
              ("glu_white", 8, 2160, 3840, torch.float32, "4k_f32")]
if what in ("f16", "all"): cases += [("glu", 32, 1080, 1920, torch.float16, "1080p_f16")]
for (kind, B, H, W, dt, tag) in cases:
    x = natural_frames(8, H, W, seed=1, dtype=dt, device=dev).repeat(B // 8, 1, 1, 1).contiguous()
    out = torch.empty_like(x)
    if kind == "glu_flat":
        x = torch.full_like(x, 0.4)
    if kind == "glu_white":
        x = torch.rand(x.shape, device=dev).to(dt)
    if kind.startswith("glu"):
        g = dict(G, intensity=1e-6) if kind in ("glu_lowgrain", "glu_flat") else G
        c = pkg.chain.PostChain(grain=g, lut=dict(lut_data=lut, strength=10.0), stencil=S, device=dev)
    else:
        ref_sums = ops.lab_moments(natural_frames(1, H, W, seed=9, dtype=dt, device=dev))
        c = pkg.chain.PostChain(grain=G, colormatch=dict(ref_sums=ref_sums, strength=1.0), lut=dict(lut_data=lut, strength=10.0), stencil=S, device=dev)
    ms, clocks = timeit(lambda: c(x, out=out))
    print(json.dumps({"lib": os.path.basename(os.environ.get("VRGDG_B200_LIB", "stock")), "row": kind + "/" + tag, "frames": B, "ms": round(ms, 4),
                      "GPx/s": round(B * H * W / ms / 1e6, 2), "checksum": float(out.float().sum().item()), "clocks": clocks}), flush=True)
    del x, out, c
    torch.cuda.empty_cache()
'''


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "f32"
    libs = [None] + sorted(glob.glob(os.path.join(PKG, "lib", "variants", "*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["VRGDG_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "what": what}], env=env, capture_output=True, text=True, timeout=600)
        print(r.stdout.strip() or ("FAILED %s: %s" % (lib, r.stderr[-600:])), flush=True)


if __name__ == "__main__":
    main()
