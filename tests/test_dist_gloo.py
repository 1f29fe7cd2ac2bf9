"""N>1 host logic on CPU: world_size-2 gloo process group (127.0.0.1).  The only collective of the path is the
all-gather of the colour-match reference sums; frames shard contiguously with no exchange."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_moments(image, row0, rows):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vrgdg_oracle as oracle
    return oracle.lab_moments_f64(image[:, row0:row0 + rows])


def oracle_counts(image):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vrgdg_oracle as oracle
    return oracle.hist_counts(image).to(torch.int32)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = importlib.import_module(PKG_NAME + ".dist")
        from helpers import natural_frames
        ref = natural_frames(1, 37, 52, seed=3)            # odd height: ranks own different row counts
        sums = d.reference_sums_distributed(ref, moments_fn=_cpu_moments)
        tiny = natural_frames(1, 1, 9, seed=4)             # fewer rows than ranks: one rank contributes zeros
        sums_tiny = d.reference_sums_distributed(tiny, moments_fn=_cpu_moments)
        hist = d.reference_histogram_distributed(ref, counts_fn=lambda img, r0, n: oracle_counts(img[:, r0:r0 + n]))
        clip = natural_frames(5, 6, 8, seed=9)                # temporal stencil: one halo frame per shard boundary
        a, b = d.shard_range(5, rank, world)
        prev, nxt = d.exchange_halo_frames(clip[a:b])
        torch.save({"sums": sums, "tiny": sums_tiny, "range": d.shard_range(11, rank, world), "prev": prev, "next": nxt, "span": (a, b), "hist": hist},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_reference_sums_allgather_two_ranks(tmp_path, oracle):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world))
    from helpers import natural_frames
    ref, tiny = natural_frames(1, 37, 52, seed=3), natural_frames(1, 1, 9, seed=4)
    assert torch.equal(r0["sums"], r1["sums"]) and torch.equal(r0["tiny"], r1["tiny"])          # identical on every rank
    assert torch.allclose(r0["sums"], oracle.lab_moments_f64(ref), rtol=1e-7, atol=1e-6)
    assert torch.allclose(r0["tiny"], oracle.lab_moments_f64(tiny), rtol=1e-7, atol=1e-6)
    assert r0["sums"][0, 0] == 37 * 52
    assert r0["range"] == (0, 6) and r1["range"] == (6, 11)
    assert torch.equal(r0["hist"], r1["hist"]) and torch.equal(r0["hist"].long(), oracle.hist_counts(ref))   # exact integers on every rank
    clip = natural_frames(5, 6, 8, seed=9)
    assert r0["span"] == (0, 3) and r1["span"] == (3, 5)
    assert r0["prev"] is None and torch.equal(r0["next"], clip[3])          # rank 0: no predecessor, successor = rank 1's first frame
    assert torch.equal(r1["prev"], clip[2]) and r1["next"] is None


def test_shard_ranges_partition_the_batch():
    d = importlib.import_module(PKG_NAME + ".dist")
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 4, 8):
            spans = [d.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_path_needs_no_process_group(oracle):
    d = importlib.import_module(PKG_NAME + ".dist")
    from helpers import natural_frames
    ref = natural_frames(1, 20, 24, seed=6)
    assert torch.equal(d.reference_sums_distributed(ref, moments_fn=_cpu_moments), oracle.lab_moments_f64(ref))
