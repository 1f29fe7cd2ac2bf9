"""bench.py — headline benchmark of the post-processing hot path (driver contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload = BASELINE.json configs[1]: fused grain + 33^3 .cube LUT + unsharp on 64 x 1920 x 1080 fp16 frames per GPU
(weak scaling: every rank owns 64 frames; frames are independent, no data-path collective).
One "step" = one pass of the fused chain over the rank's 64 frames = ONE kernel launch (k_tile).
  value : MP/s, frames resident in HBM (CUDA events, max over ranks)
  e2e   : MP/s through the public API (PostChain.run_host) from pinned HOST frames to pinned HOST frames, H2D and
          D2H copies inside the timed region
  --impl reference : the reference's CPU path for the same chain (oracle port of the reference nodes; /root/reference
          does not exist on the GPU box) on all host cores, rank 0 only, bounded sample per step
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

def host_threads():
    """Threads for the CPU arm: one per PHYSICAL core (torch's own default; measured on the GPU box: 128 SMT threads run the
    reference chain 3x slower than 64)."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
    except Exception:
        n = None
    return int(n or max(1, (os.cpu_count() or 2) // 2))


if "reference" in sys.argv:      # the CPU arm uses every core; torchrun would otherwise pin OMP_NUM_THREADS=1
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    os.environ["MKL_NUM_THREADS"] = str(host_threads())

import torch  # noqa: E402

PKG = "comfyui-vrgamedevgirl_b200"
LUT_FILE = os.path.join(ROOT, PKG, "LUTS", "B200 Vintage 33.cube")
FRAMES, H, W = 64, 1080, 1920
GRAIN = dict(intensity=0.04, saturation_mix=0.5, seed=42)
SHARPEN = 0.5
METRIC = "megapixels/sec"
WORKLOAD = "configs[1]: fused grain + 33^3 .cube LUT + unsharp, 64x1920x1080 fp16 frames per GPU"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("dram_bytes_per_launch"), d
    return None, None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled every 10 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_chain_sample(frames, steps=1, threads=None):
    """The reference's CPU path for the same chain (oracle port, fp32 as users run it) on `frames` 1080p frames.
    Returns (MP/s, seconds, threads)."""
    import vrgdg_oracle as oracle
    from helpers import natural_frames
    if threads:
        torch.set_num_threads(threads)
    x = natural_frames(frames, H, W, seed=0)
    lut = oracle.parse_cube(LUT_FILE)
    best = None
    for _ in range(steps):
        torch.manual_seed(123)
        t0 = time.perf_counter()
        a = oracle.film_grain(x, GRAIN["intensity"], GRAIN["saturation_mix"], batch_size=4)
        b = oracle.apply_lut(a, lut, 10.0)
        oracle.unsharp_numpy(b, SHARPEN)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames * H * W / 1e6 / best, best, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = host_threads()
    torch.set_num_threads(threads)
    sample_frames = 4
    import vrgdg_oracle as oracle
    from helpers import natural_frames
    x = natural_frames(sample_frames, H, W, seed=0)
    lut = oracle.parse_cube(LUT_FILE)

    def step():
        a = oracle.film_grain(x, GRAIN["intensity"], GRAIN["saturation_mix"], batch_size=4)
        b = oracle.apply_lut(a, lut, 10.0)
        return oracle.unsharp_numpy(b, SHARPEN)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    mps = sample_frames * H * W / 1e6 / dt
    sample = "%d x 1080p fp32 frames per step (reference CPU nodes FastFilmGrain -> VRGDG_LUTS -> FastUnsharpSharpen numpy path; fp16 on CPU is not what users run)" % sample_frames
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(mps, 3), "unit": "MP/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": round(mps, 3), "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(mps, 3), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pkg = importlib.import_module(PKG)
    nv = pkg._native
    nv.load_library()
    from helpers import natural_frames

    # synthetic clip, generated on the device; rank r owns absolute frames [r*FRAMES, (r+1)*FRAMES)
    base = natural_frames(8, H, W, seed=rank, dtype=torch.float16, device=dev)
    x = base.repeat(FRAMES // 8, 1, 1, 1).contiguous()
    x += (torch.rand(FRAMES, 1, 1, 1, device=dev) * 0.02).half()
    x.clamp_(0, 1)
    del base
    out = torch.empty_like(x)
    lut = pkg.VRGDG_LUTS._parse_cube_file(LUT_FILE)
    chain = pkg.chain.PostChain(grain=GRAIN, lut=dict(lut_data=lut, strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=SHARPEN, border=nv.BORDER_REPLICATE), device=dev)
    first = rank * FRAMES
    npix = FRAMES * H * W

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        chain(x, first_frame=first, out=out)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = nv.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        chain(x, first_frame=first, out=out)
    e1.record()
    barrier()
    launches = nv.launch_count() - launches0
    sampler.stop_flag.set()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * npix / 1e6 / (ms_step / 1e3)
    tile_path = nv.last_tile_path()

    # ---- end to end through the public API: pinned host frames -> pinned host frames ----
    host_in = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    host_in.copy_(x)
    e2e_steps = max(1, min(args.steps, 5))
    host_out = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)      # pinned once, reused every step
    chain.run_host(host_in, chunk_frames=8, first_frame=first, out=host_out)      # warm-up
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for _ in range(e2e_steps):
        res = chain.run_host(host_in, chunk_frames=8, first_frame=first, out=host_out)
        checksum += float(res[0, 0, 0, 0])          # device->host result is read on the host
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = world * npix / 1e6 / (e2e_ms / 1e3)
    frame_bytes = x.numel() * x.element_size()

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = npix * 12                      # 6 B read + 6 B written per pixel (fp16 RGB), SURVEY 8(d)
        achieved = alg_bytes / (ms_step / 1e3) / 1e9
        traffic, prof = recorded_traffic()
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": FRAMES, "height": H, "width": W, "frame_dtype": "f16", "lut": os.path.basename(LUT_FILE),
                       "distribution": "natural-like (4 octaves of upsampled noise + 2% white)", "parallelism": "frame-sharded dp%d" % world,
                       "l2": "input (796 MB per GPU) and output are each larger than L2 (126 MB); no flush needed", "tile_path": tile_path},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "kernel": "k_tile<half, grain|lut, unsharp>", "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": peak_src},
            "e2e": {"value": round(e2e_value, 1), "unit": "MP/s", "h2d_bytes_per_step": frame_bytes, "d2h_bytes_per_step": frame_bytes,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps, "api": "PostChain.run_host(pinned frames, chunk_frames=8)"},
            "gpu_launches": launches,
            "clocks": sampler.result(),
        }
        if world == 1 and not args.no_cpu:
            mps, secs, threads = cpu_chain_sample(2, steps=2)
            line["cpu_baseline"] = {"value": round(mps, 3), "unit": "MP/s", "cores": threads, "kind": "port",
                                    "sample": "2 x 1080p fp32 frames, best of 2, oracle port of FastFilmGrain -> VRGDG_LUTS -> FastUnsharpSharpen (%.1f s)" % secs}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
        run_b200(args)


if __name__ == "__main__":
    main()
