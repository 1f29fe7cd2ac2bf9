"""Where the time of an unchanged ComfyUI workflow goes (GPU box only): the four stock node classes on pageable host tensors,
per-node wall clock + cProfile of one step.   python tools/stock_profile.py [frames]"""
import cProfile
import importlib
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import natural_frames  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
hx = natural_frames(n, 2160, 3840, seed=1, device=dev).cpu()
hr = natural_frames(1, 2160, 3840, seed=9, device=dev).cpu()
nodes = (pkg.FastFilmGrain(), pkg.ColorMatchToReference(), pkg.VRGDG_LUTS(), pkg.FastUnsharpSharpen())


def step(times=None):
    def tick(name, t0):
        torch.cuda.synchronize()
        if times is not None:
            times.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    t = time.perf_counter(); a = nodes[0].apply_grain(hx, 0.04, 0.5, 4)[0]; tick("grain", t)
    t = time.perf_counter(); b = nodes[1].match_color(a, hr, 1.0, 1)[0]; tick("colormatch", t)
    t = time.perf_counter(); c = nodes[2].apply_lut(b, "B200 Vintage 33.cube", "auto", 10.0)[0]; tick("lut", t)
    t = time.perf_counter(); d = nodes[3].apply_unsharp(c, 0.5, False)[0]; tick("unsharp", t)
    return d


step(); step()
times = {}
t0 = time.perf_counter()
for _ in range(3):
    step(times)
total = (time.perf_counter() - t0) / 3 * 1e3
print("frames", n, "ms per step %.1f" % total, {k: [round(x, 1) for x in v] for k, v in times.items()})
print("pinned?", [nodes[0].apply_grain(hx, 0.04, 0.5, 4)[0].is_pinned()])
pr = cProfile.Profile()
pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
