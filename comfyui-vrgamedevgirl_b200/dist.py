"""Multi-GPU: one process per GPU (torchrun), frames sharded contiguously along the batch axis.

Every frame is independent for grain (keyed by the absolute frame index), the LUT, the stencils and colour match
(moments are per frame, nodes.py:109-110), so shards exchange nothing on the data path.  The single collective is
for colour match only: the REFERENCE image's LAB sums (nodes.py:98-100).  Its rows are sharded across ranks, each
rank reduces its rows to 7 doubles, one all-gather (56 bytes per rank) shares them, and every rank adds them in
rank order, so all ranks hold bit-identical reference statistics.
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world_size):
    """Contiguous split of [0, n_frames): rank r owns [start, stop); earlier ranks take the remainder."""
    base, rem = divmod(int(n_frames), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def row_range(height, rank, world_size):
    return shard_range(height, rank, world_size)


def reference_sums_distributed(reference_image, moments_fn=None, group=None):
    """LAB raw sums [1,7] float64 of a [1,H,W,3] reference image, computed cooperatively.

    moments_fn(image, row0, rows) -> [1,7] float64 on the communication device; defaults to the CUDA kernel.
    Ranks that own no rows contribute zeros."""
    if moments_fn is None:
        from . import ops
        moments_fn = lambda img, r0, n: ops.lab_moments(img, r0, n)
    if not (dist.is_available() and dist.is_initialized()):
        return moments_fn(reference_image, 0, int(reference_image.shape[1]))
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    r0, r1 = row_range(int(reference_image.shape[1]), rank, world)
    if r1 > r0:
        mine = moments_fn(reference_image, r0, r1 - r0).reshape(1, 7).to(torch.float64)
    else:
        mine = torch.zeros(1, 7, dtype=torch.float64, device=reference_image.device if reference_image.device.type == "cuda" else "cpu")
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine.contiguous(), group=group)
    total = torch.zeros_like(mine)
    for part in gathered:            # fixed rank order -> identical on every rank
        total += part
    return total


def reference_histogram_distributed(reference_image, counts_fn=None, group=None):
    """256-bin RGB counts [1,3,256] int32 of a [1,H,W,3] reference image for the histogram colour-match mode (an extension of this
    package; BASELINE.json: "NCCL all-gather for the global reference histogram"): rows sharded like reference_sums_distributed,
    one all-gather of 3 x 256 counters per rank, summed.  Integer counts: identical on every rank and for every world size."""
    if counts_fn is None:
        from . import ops
        counts_fn = lambda img, r0, n: ops.hist_counts(img, r0, n)
    H = int(reference_image.shape[1])
    if not (dist.is_available() and dist.is_initialized()):
        return counts_fn(reference_image, 0, H)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    r0, r1 = row_range(H, rank, world)
    if r1 > r0:
        mine = counts_fn(reference_image, r0, r1 - r0).reshape(1, 3, 256).to(torch.int32).contiguous()
    else:
        mine = torch.zeros(1, 3, 256, dtype=torch.int32, device=reference_image.device if reference_image.device.type == "cuda" else "cpu")
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    total = torch.zeros_like(mine)
    for part in gathered:
        total += part
    return total


def exchange_halo_frames(frames, group=None):
    """(prev, next) = the last frame of the previous rank and the first frame of the next rank ([H,W,3] each, None at the clip's
    ends) for the temporal 3-frame stencil (configs[4], an extension without a reference counterpart): one frame sent to each
    neighbour, point to point.  Without a process group: (None, None)."""
    if not (dist.is_available() and dist.is_initialized()):
        return None, None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1 or frames.shape[0] == 0:
        return None, None
    first, last = frames[0].contiguous(), frames[-1].contiguous()
    prev = torch.empty_like(first) if rank > 0 else None
    nxt = torch.empty_like(first) if rank + 1 < world else None
    ops_ = []
    if rank > 0:
        ops_ += [dist.P2POp(dist.isend, first, rank - 1, group), dist.P2POp(dist.irecv, prev, rank - 1, group)]
    if rank + 1 < world:
        ops_ += [dist.P2POp(dist.isend, last, rank + 1, group), dist.P2POp(dist.irecv, nxt, rank + 1, group)]
    for req in dist.batch_isend_irecv(ops_):
        req.wait()
    return prev, nxt
