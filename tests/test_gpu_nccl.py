"""The path's one collective ON THE GPUs: colour match's reference-image statistics (nodes.py:98-100) with the reference rows
sharded over the ranks, k_lab_moments on each rank's rows, ONE all-gather of 7 doubles per rank over NCCL, rank-order fold.

world_size = min(2, visible GPUs): NCCL refuses two ranks on one device, so a single-GPU box runs the same code with world_size 1
(the all-gather still goes through NCCL).  bench.py --gpus N runs the identical call inside every timed step."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pkg = importlib.import_module(PKG_NAME)
        d = importlib.import_module(PKG_NAME + ".dist")
        from helpers import natural_frames
        ref = natural_frames(1, 270, 480, seed=3).to(dev)            # 270 rows: ranks own 135 rows each; vector path (W % 4 == 0)
        before = pkg._native.launch_count()
        sums = d.reference_sums_distributed(ref)
        launched = pkg._native.launch_count() - before
        odd = natural_frames(1, 37, 53, seed=4).to(dev)              # odd height and width: unequal row counts, scalar path
        sums_odd = d.reference_sums_distributed(odd)
        hist = d.reference_histogram_distributed(ref)                # histogram mode: 3 x 256 counters per rank, same collective shape
        clip = natural_frames(3 * world, 24, 32, seed=7).to(dev)     # temporal extension: one halo frame per shard boundary
        ca, cb = d.shard_range(3 * world, rank, world)
        prev, nxt = d.exchange_halo_frames(clip[ca:cb])
        temporal = pkg.ops.temporal_sharpen(clip[ca:cb].contiguous(), 0.7, prev, nxt)
        tiny = natural_frames(1, 1, 8, seed=5).to(dev)               # fewer rows than ranks: a rank contributes zeros
        sums_tiny = d.reference_sums_distributed(tiny)
        # the statistics drive a sharded colour match: rank r owns frames [r*2, r*2+2) of a 2*world-frame clip
        frames = natural_frames(2 * world, 64, 96, seed=6)
        a, b = d.shard_range(2 * world, rank, world)
        chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), colormatch=dict(ref_sums=sums, strength=1.0), device=dev)
        shard_out = chain(frames[a:b].to(dev), first_frame=a)
        torch.cuda.synchronize(dev)
        torch.save({"sums": sums.cpu(), "odd": sums_odd.cpu(), "tiny": sums_tiny.cpu(), "launched": launched, "shard": shard_out.cpu(), "range": (a, b), "hist": hist.cpu(), "temporal": temporal.cpu(), "trange": (ca, cb),
                    "backend": dist.get_backend()}, os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_reference_moments_allgather_over_nccl(pkg, cuda_device, tmp_path):
    world = min(2, torch.cuda.device_count())
    assert world >= 1
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    d = importlib.import_module(PKG_NAME + ".dist")
    from helpers import natural_frames
    assert all(r["backend"] == "nccl" and r["launched"] >= 2 for r in res)      # k_lab_moments + k_moments_final ran on every rank
    for key, (seed, H, W) in (("sums", (3, 270, 480)), ("odd", (4, 37, 53)), ("tiny", (5, 1, 8))):
        for r in res[1:]:
            assert torch.equal(r[key], res[0][key]), key                        # bit-identical on every rank
        img = natural_frames(1, H, W, seed=seed).to(cuda_device)
        # single-GPU evaluation over the same row partition, folded in rank order: bit-identical
        parts = torch.zeros(1, 7, dtype=torch.float64, device=cuda_device)
        for rk in range(world):
            r0, r1 = d.row_range(H, rk, world)
            if r1 > r0:
                parts += pkg.ops.lab_moments(img, r0, r1 - r0)
        assert torch.equal(res[0][key], parts.cpu()), key
        # and equal to the whole-frame single-GPU statistics to fp64 rounding
        whole = pkg.ops.lab_moments(img).cpu()
        assert torch.allclose(res[0][key], whole, rtol=1e-12, atol=1e-9), key
        assert res[0][key][0, 0] == H * W
    # histogram counts: exact integers, identical on every rank, equal to the single-GPU counts
    img = natural_frames(1, 270, 480, seed=3).to(cuda_device)
    for r in res:
        assert torch.equal(r["hist"], pkg.ops.hist_counts(img).cpu())
    # temporal sharpen over shards with exchanged halo frames == the whole clip on one GPU
    clip = natural_frames(3 * world, 24, 32, seed=7).to(cuda_device)
    whole_t = pkg.ops.temporal_sharpen(clip, 0.7).cpu()
    for r in res:
        ca, cb = r["trange"]
        assert torch.equal(r["temporal"], whole_t[ca:cb])
    # the sharded colour match equals the single-GPU run on the whole clip
    frames = natural_frames(2 * world, 64, 96, seed=6).to(cuda_device)
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), colormatch=dict(ref_sums=res[0]["sums"], strength=1.0), device=cuda_device)
    whole = chain(frames, first_frame=0).cpu()
    for r in res:
        a, b = r["range"]
        assert torch.equal(r["shard"], whole[a:b])
