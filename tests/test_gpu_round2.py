"""Round-2 GPU parity tests (through the C ABI):
  * 64^3 / 65^3 LUTs (the common sizes of the reference's own files) bit-exact against outputs of the reference itself
  * fp16 / bf16 frames against THE ORACLE run on the up-cast input (not against our own fp32 kernels): grain, unsharp, LUT and the
    configs[1] chain in the benchmarked arithmetic (fast_math), at 1080p; tolerance 1 ulp of the 16-bit type
  * the full configs[3] chain through PostChain on external noise (statistics and apply see the same grained frames)
  * host-side contracts added this round (`out` validation, chunking of CUDA inputs)"""
import os

import numpy as np
import pytest
import torch

from helpers import LUTS, load_golden, natural_frames, t, write_big_cube

pytestmark = pytest.mark.gpu
TOL = 1e-5
ULP = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}      # spacing of the 16-bit type in [0.5, 1)


def maxdiff(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def _lut33(pkg):
    return pkg.VRGDG_LUTS._parse_cube_file(os.path.join(LUTS, "B200 Vintage 33.cube"))


# ------------------------------------------------------------------------------------------------------
# LUT sizes 64 and 65
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [64, 65])
def test_lut_64_65_bit_exact_vs_reference_outputs(pkg, cuda_device, tmp_path, size):
    g = load_golden("lut_big")
    data = pkg.VRGDG_LUTS._parse_cube_file(write_big_cube(str(tmp_path / ("big_%d.cube" % size)), size))
    assert data["size"] == size
    x, xn = t(g["x"]).to(cuda_device), t(g["xn"]).to(cuda_device)
    before = pkg._native.launch_count()
    out = pkg.VRGDG_LUTS._apply_cube_lut(x, data["lut"], data["domain_min"], data["domain_max"])
    assert pkg._native.launch_count() > before
    assert torch.equal(out.cpu(), t(g["s%d__s10" % size]))
    lut_nodes = __import__("importlib").import_module(pkg.__name__ + ".lut_nodes")
    assert torch.equal(lut_nodes._run_lut(t(g["x"]), data, 3.5), t(g["s%d__s3p5" % size]))            # strength blend, host tensor in / out
    assert torch.equal(lut_nodes._run_lut(xn, data, 10.0).cpu(), t(g["s%d__nat" % size]))
    o16 = lut_nodes._run_lut(t(g["x"]).half(), data, 10.0)
    assert o16.dtype == torch.float16 and torch.equal(o16, t(g["s%d__fp16" % size]))                     # one rounding of identical fp32 values
    # uint8 BGR frames through the same table == decode -> LUT -> encode of the oracle
    # fused chain with the big table: fused == separate kernels (the cell table leaves L1 at this size: 25-26 MB)
    nv = pkg._native
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=7), lut=dict(lut_data=data, strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=cuda_device)
    xf = natural_frames(2, 136, 248, seed=size, device=cuda_device)
    a = pkg.ops.grain(xf, 0.04, 0.5, 0.5, seed=7, frame0=3)
    b = pkg.ops.lut3d_apply(a, data["lut"].to(cuda_device), [0, 0, 0], [1, 1, 1], 1.0, 0.0)
    c = pkg.ops.stencil3x3(b, nv.STENCIL_BOX_UNSHARP, 0.5, nv.BORDER_REPLICATE)
    assert maxdiff(chain(xf, first_frame=3), c) <= 2e-6


def test_lut_65_full_size_vs_oracle(pkg, cuda_device, oracle, tmp_path):
    """a whole 1080p frame through the 65^3 table (26 MB cell table, L2-resident gather) equals the oracle bit for bit"""
    path = write_big_cube(str(tmp_path / "big_65.cube"), 65)
    data = pkg.VRGDG_LUTS._parse_cube_file(path)
    x = natural_frames(1, 1080, 1920, seed=65)
    out = pkg.VRGDG_LUTS._apply_cube_lut(x.to(cuda_device), data["lut"], data["domain_min"], data["domain_max"]).cpu()
    assert torch.equal(out, oracle.apply_lut(x, oracle.parse_cube(path), 10.0))


# ------------------------------------------------------------------------------------------------------
# polynomial LUT cells of the fast chains (vrgdg_math.cuh "polynomial cells")
# ------------------------------------------------------------------------------------------------------
def _fast_chain_vs_oracle(pkg, oracle, cuda_device, lut_data, olut, x, z, strength=10.0):
    nv = pkg._native
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=0), lut=dict(lut_data=lut_data, strength=strength),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5, border=nv.BORDER_REPLICATE), device=cuda_device)
    ref = oracle.chain_grain_lut_unsharp(x, z, 0.04, 0.5, olut, strength, 0.5)
    fast = chain(x.to(cuda_device), ext_noise=z.to(cuda_device), fast_math=True)        # contracted arithmetic + polynomial cells
    exact = chain(x.to(cuda_device), ext_noise=z.to(cuda_device))                        # reference op sequence + corner cells
    return maxdiff(fast, ref), maxdiff(exact, ref)


SHIPPED = {17: "Warm_Fade_17.cube", 25: "Teal_Orange_25.cube", 32: "Bleach_Bypass_32.cube", 33: "B200 Vintage 33.cube"}


@pytest.mark.parametrize("size", [17, 25, 32, 33, 64, 65])
def test_polynomial_lut_cells_fast_chain_vs_oracle(pkg, cuda_device, oracle, tmp_path, size):
    """the benchmarked arithmetic (fast_math: the lookup evaluates the trilinear polynomial from its coefficient cells) against the
    oracle composition on the reference's noise, for every table size the reference ships (25, 32, 33, 64, 65) and 17"""
    x = natural_frames(2, 270, 480, seed=size)
    z = torch.randn(x.shape, generator=torch.Generator().manual_seed(size + 1))
    if size in SHIPPED:
        path = os.path.join(LUTS, SHIPPED[size])
    else:
        path = write_big_cube(str(tmp_path / ("big_%d.cube" % size)), size)
    fast, exact = _fast_chain_vs_oracle(pkg, oracle, cuda_device, pkg.VRGDG_LUTS._parse_cube_file(path), oracle.parse_cube(path), x, z)
    assert exact <= 2e-6 and fast <= 4e-6, (fast, exact)
    fast, exact = _fast_chain_vs_oracle(pkg, oracle, cuda_device, pkg.VRGDG_LUTS._parse_cube_file(path), oracle.parse_cube(path), x, z, strength=3.5)
    assert exact <= 2e-6 and fast <= 4e-6, (fast, exact)


def test_polynomial_lut_cells_whole_4k_frame_fast_equals_exact(pkg, cuda_device):
    """BASELINE size: one whole 3840x2160 fp32 frame through the fused chain on the same external noise, benchmarked arithmetic
    (coefficient cells, contracted FMAs, fast stencil) against the reference op sequence (corner cells), which the small-size tests
    pin to the oracle; every pixel within 4e-6"""
    nv = pkg._native
    x = natural_frames(1, 2160, 3840, seed=4, device=cuda_device)
    z = torch.randn(x.shape, generator=torch.Generator(device=cuda_device).manual_seed(5), device=cuda_device)
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=0), lut=dict(lut_data=_lut33(pkg), strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5, border=nv.BORDER_REPLICATE), device=cuda_device)
    fast, exact = chain(x, ext_noise=z, fast_math=True), chain(x, ext_noise=z)
    assert nv.last_tile_path() == "tma"
    assert float((fast - exact).abs().max()) <= 4e-6
    assert float(fast.min()) >= 0.0 and float(fast.max()) <= 1.0


def test_polynomial_lut_cells_any_table_range(pkg, cuda_device, oracle):
    """tables whose values leave [0,1], a constant channel and a non-unit input domain through the coefficient cells; the packed
    buffer holds both tables (corner cells for the exact entry points, coefficient cells for the fast chains)"""
    S = 17
    g = torch.Generator().manual_seed(5)
    ax = torch.linspace(0, 1, S)
    bb, gg, rr = torch.meshgrid(ax, ax, ax, indexing="ij")
    lut = torch.stack([rr * 2.5 - 0.75 + 0.05 * torch.rand(rr.shape, generator=g),          # range ~[-0.75, 1.8]
                       torch.full_like(gg, 0.25),                                              # constant channel
                       0.4 + 0.2 * bb * gg], dim=-1).float().contiguous()                      # narrow range
    data = dict(lut=lut, size=S, domain_min=torch.tensor([0.1, 0.0, 0.05]), domain_max=torch.tensor([0.9, 1.0, 0.8]), title="range")
    x = natural_frames(1, 128, 192, seed=7)
    z = torch.randn(x.shape, generator=torch.Generator().manual_seed(8))
    fast, exact = _fast_chain_vs_oracle(pkg, oracle, cuda_device, data, data, x, z)
    assert exact <= 2e-6 and fast <= 4e-6, (fast, exact)
    packed = pkg.ops.pack_lut(lut, cuda_device)
    n = S ** 3
    assert packed.data.numel() == n * 48
    corner, poly = packed.data[:n * 24].reshape(n, 3, 8).cpu(), packed.data[n * 24:].reshape(n, 3, 8).cpu()
    assert torch.equal(poly[:, :, 0], corner[:, :, 0])                                         # k000 = c000
    assert torch.equal(poly[:, 1, 1:], torch.zeros(n, 7))                                      # constant channel: no gradient terms
    want = (corner[:, :, 7].double() - corner[:, :, 6].double() - corner[:, :, 5].double() - corner[:, :, 3].double()
            + corner[:, :, 4].double() + corner[:, :, 2].double() + corner[:, :, 1].double() - corner[:, :, 0].double())
    assert torch.equal(poly[:, :, 7], want.float())                                            # k111, formed in double, rounded once


# ------------------------------------------------------------------------------------------------------
# 16-bit frames against the oracle on the up-cast input
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_half_precision_frames_vs_oracle_on_upcast_input_1080p(pkg, cuda_device, oracle, dt):
    nv = pkg._native
    x16 = natural_frames(1, 1080, 1920, seed=16).to(dt)
    z16 = torch.randn(x16.shape, generator=torch.Generator().manual_seed(17)).to(dt)
    x32, z32 = x16.float(), z16.float()
    xd, zd = x16.to(cuda_device), z16.to(cuda_device)
    lut = _lut33(pkg)
    olut = oracle.parse_cube(os.path.join(LUTS, "B200 Vintage 33.cube"))
    ulp = ULP[dt]

    # grain on external noise: the kernel reproduces the fp32 op sequence, then rounds once
    ref = oracle.film_grain(x32, 0.04, 0.5, 0, noise=z32)
    got = pkg.ops.grain(xd, 0.04, 0.5, 0.5, seed=0, ext_noise=zd)
    assert got.dtype == dt and torch.equal(got.cpu(), ref.to(dt))

    # unsharp (NumPy-path semantics): fast arithmetic for 16-bit frames -> within one spacing of the rounded oracle
    ref = oracle.unsharp_numpy(x32, 0.5)
    got = pkg.ops.stencil3x3(xd, nv.STENCIL_BOX_UNSHARP, 0.5, nv.BORDER_REPLICATE)
    assert got.dtype == dt and maxdiff(got.float(), ref) <= ulp

    # LUT: exact fp32 lookup, one rounding
    ref = oracle.apply_lut(x32, olut, 10.0)
    got = pkg.ops.lut3d_apply(xd, lut["lut"].to(cuda_device), [0, 0, 0], [1, 1, 1], 1.0, 0.0)
    assert torch.equal(got.cpu(), ref.to(dt))

    # configs[1] chain, the arithmetic bench.py runs (FMA-contracted, fast stencil), on the reference's noise
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=0), lut=dict(lut_data=lut, strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5, border=nv.BORDER_REPLICATE), device=cuda_device)
    ref = oracle.chain_grain_lut_unsharp(x32, z32, 0.04, 0.5, olut, 10.0, 0.5)
    fast = chain(xd, ext_noise=zd, fast_math=True)
    assert nv.last_tile_path() == "tma" and fast.dtype == dt
    assert maxdiff(fast.float(), ref) <= ulp
    exact = chain(xd, ext_noise=zd)
    assert maxdiff(exact.float(), ref) <= ulp


# ------------------------------------------------------------------------------------------------------
# configs[3] chain through the public class, statistics and apply on the same external noise
# ------------------------------------------------------------------------------------------------------
def test_postchain_full_chain_on_reference_noise(pkg, cuda_device):
    nv = pkg._native
    g = load_golden("chain")
    x, z = t(g["x"]).to(cuda_device), t(g["z"]).to(cuda_device)
    chain = pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=0), colormatch=dict(reference_image=t(g["ref"]), strength=1.0),
                                lut=dict(lut_data=_lut33(pkg), strength=10.0),
                                stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5, border=nv.BORDER_REPLICATE), device=cuda_device)
    out = chain(x, ext_noise=z)
    assert maxdiff(out, t(g["grain_cm_lut_unsharp"])) <= TOL
    fast = chain(x, ext_noise=z, fast_math=True)                       # the arithmetic the benchmark runs
    assert maxdiff(fast, t(g["grain_cm_lut_unsharp"])) <= TOL
    # the moments pass honours the external noise: equal to moments of the materialised grained frames
    grained = pkg.ops.grain(x, 0.04, 0.5, 0.5, seed=0, ext_noise=z)
    d = nv.ChainDesc()
    d.grain_enabled, d.grain_intensity, d.grain_sat, d.grain_one_minus_sat = 1, 0.04, 0.5, 0.5
    assert torch.equal(pkg.ops.chain_lab_moments(x, d, ext_noise=z), pkg.ops.lab_moments(grained))


def test_one_call_colormatch_chain_schedules_agree(pkg, cuda_device):
    """vrgdg_chain_cm_apply: f-plane schedule (fp32), recompute schedule, any group size, and the three-call path give the same frames"""
    nv = pkg._native
    lut = _lut33(pkg)
    ref = natural_frames(1, 50, 60, seed=21) * 0.8
    for dt, W in ((torch.float32, 96), (torch.float32, 93), (torch.float16, 96), (torch.uint8, 96)):
        x = natural_frames(5, 72, W, seed=20)
        x = (x * 255).to(torch.uint8) if dt == torch.uint8 else x.to(dt)
        x = x.to(cuda_device)
        refd = (ref * 255).to(torch.uint8) if dt == torch.uint8 else ref.to(dt)
        mk = lambda: pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), colormatch=dict(reference_image=refd, strength=0.8),
                                         lut=dict(lut_data=lut, strength=10.0), stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=cuda_device)
        base = mk()
        before = nv.launch_count()
        fused = base(x, first_frame=7)
        assert nv.launch_count() - before >= 4
        three = mk()
        three.split = True
        want = three(x, first_frame=7)
        rec = mk()
        rec.recompute = True
        assert torch.equal(rec(x, first_frame=7), want), dt                      # same kernels, grouped: bit-identical
        tol = 0 if dt != torch.float32 else 2e-6                                # fp32: the f-plane pass (other dtypes always recompute)
        assert maxdiff(fused.float(), want.float()) <= tol, (dt, W)
        for g in (1, 2, 3):
            c = mk()
            c.group_frames = g
            assert torch.equal(c(x, first_frame=7), fused), (dt, g)              # group size never changes a pixel
            c.serial = True                                                      # groups one after the other instead of pipelined
            assert torch.equal(c(x, first_frame=7), fused), (dt, g, "serial")
        # shard invariance through the one-call path
        assert torch.equal(mk()(x[2:].contiguous(), first_frame=9), fused[2:])
    # colour match without LUT / stencil (streaming second pass) and one reference per frame (n_ref == B)
    x = natural_frames(4, 40, 64, seed=22, device=cuda_device)
    refs = natural_frames(4, 30, 44, seed=23, device=cuda_device)
    ref_sums = pkg.ops.lab_moments(refs)
    d = nv.ChainDesc()
    d.colormatch_enabled, d.cm_t, d.cm_one_minus_t = 1, 1.0, 0.0
    got, _ = pkg.ops.chain_cm_apply(x, d, ref_sums, group_frames=3)
    want = pkg.ops.colormatch_apply(x, pkg.ops.colormatch_params(pkg.ops.lab_moments(x), ref_sums), 1.0, 0.0)
    assert maxdiff(got, want) <= 2e-6
    got_r, _ = pkg.ops.chain_cm_apply(x, d, ref_sums, recompute=True)
    assert torch.equal(got_r, want)
    with pytest.raises(ValueError):
        pkg.ops.chain_cm_apply(x, nv.ChainDesc(), ref_sums)                      # no colour-match stage in the descriptor
    with pytest.raises(ValueError):
        pkg.ops.chain_cm_apply(x, d, ref_sums[:3])                               # reference batch neither 1 nor B


def test_pipelined_colormatch_schedule_many_groups(pkg, cuda_device):
    """the two-stream schedule (statistics pass of group g+1 under the apply pass of group g, double-buffered f-planes) on enough
    groups to wrap the buffers several times, back to back on the caller's stream, against the serial schedule"""
    nv = pkg._native
    x = natural_frames(11, 136, 248, seed=30, device=cuda_device)
    ref = natural_frames(1, 60, 80, seed=31, device=cuda_device)
    mk = lambda: pkg.chain.PostChain(grain=dict(intensity=0.04, saturation_mix=0.5, seed=42), colormatch=dict(reference_image=ref, strength=1.0),
                                     lut=dict(lut_data=_lut33(pkg), strength=10.0), stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=cuda_device)
    serial = mk()
    serial.group_frames, serial.serial = 2, True
    want = serial(x, first_frame=3)
    piped = mk()
    piped.group_frames = 2                                      # 6 groups: buffers wrap three times
    outs = [piped(x, first_frame=3) for _ in range(4)]          # consecutive calls reuse the scratch: ordering through the caller's stream
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, want)
    # work queued on the caller's stream AFTER the call sees the finished frames (join), on a non-default stream too
    side = torch.cuda.Stream(cuda_device)
    with torch.cuda.stream(side):
        o = piped(x, first_frame=3)
        total = o.double().sum()
    side.synchronize()
    assert float(total) == float(want.double().sum())


def test_moments_vector_and_scalar_paths_agree(pkg, cuda_device, oracle):
    """W % 4 == 0 takes the 48-byte vector path, odd widths the scalar one; both against the oracle, with and without grain"""
    for W in (64, 61):
        x = natural_frames(3, 40, W, seed=W)
        sums = pkg.ops.lab_moments(x.to(cuda_device)).cpu()
        ref = oracle.lab_moments_f64(x)
        n = float(ref[0, 0])
        assert torch.equal(sums[:, 0], ref[:, 0])
        assert float((sums[:, 1:4] - ref[:, 1:4]).abs().max()) / n < 5e-5
        assert torch.allclose(sums[:, 4:7], ref[:, 4:7], rtol=5e-6, atol=0.0)
    # in-kernel grain: moments pass == moments of the frames the grain kernel writes (same generator, same arithmetic)
    nv = pkg._native
    for W, dt in ((64, torch.float32), (61, torch.float32), (64, torch.float16)):
        x = natural_frames(2, 40, W, seed=3 * W, dtype=dt, device=cuda_device)
        d = nv.ChainDesc()
        d.grain_enabled, d.grain_intensity, d.grain_sat, d.grain_one_minus_sat, d.grain_seed, d.grain_frame0 = 1, 0.04, 0.5, 0.5, 42, 5
        a = pkg.ops.chain_lab_moments(x, d)
        b = pkg.ops.lab_moments(pkg.ops.grain(x.float(), 0.04, 0.5, 0.5, seed=42, frame0=5))
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-4), (W, dt)


# ------------------------------------------------------------------------------------------------------
# host-side contracts
# ------------------------------------------------------------------------------------------------------
def test_chain_out_must_match_the_frames(pkg, cuda_device):
    nv = pkg._native
    x = natural_frames(2, 40, 64, seed=1, device=cuda_device)
    chain = pkg.chain.PostChain(stencil=dict(op=nv.STENCIL_BOX_UNSHARP, strength=0.5), device=cuda_device)
    good = torch.empty_like(x)
    assert chain(x, out=good) is good
    for bad in (torch.empty((1, 40, 64, 3), device=cuda_device), torch.empty_like(x, dtype=torch.float16), torch.empty(x.shape),
                torch.empty((2, 40, 64, 6), device=cuda_device)[..., :3], x):
        with pytest.raises(ValueError):
            chain(x, out=bad)


def test_stream_frames_chunks_cuda_inputs_too(pkg, cuda_device):
    rt = __import__("importlib").import_module(pkg.__name__ + "._runtime")
    x = natural_frames(5, 24, 32, seed=2, device=cuda_device)
    calls = []

    def fn(frames, first):
        calls.append((int(frames.shape[0]), int(first)))
        return frames * 0.5
    out = rt.stream_frames(x, fn, 2, cuda_device)
    assert calls == [(2, 0), (2, 2), (1, 4)] and torch.equal(out, x * 0.5)
    calls.clear()
    out = rt.stream_frames(x, fn, 0, torch.device("cpu"))
    assert calls == [(5, 0)] and out.device.type == "cpu"
    # host frames: more chunks than staging buffers, pinned and pageable sources
    for src in (x.cpu(), x.cpu().pin_memory()):
        calls.clear()
        out = rt.stream_frames(src, fn, 1, torch.device("cpu"), cuda_device, depth=2)
        assert [c[1] for c in calls] == [0, 1, 2, 3, 4] and torch.equal(out, src * 0.5)
        assert out.is_pinned()                                       # small host results are pinned: the next node uploads at PCIe speed


def test_stream_frames_byte_cap_and_pageable_staging(pkg, cuda_device, monkeypatch):
    """host batches are cut into pipeline chunks of at most VRGDG_STREAM_CHUNK_BYTES whatever the caller's chunk is; pageable sources
    go through the pinned staging ring; the LUT node streams host frames like the other nodes (pinned result)"""
    rt = __import__("importlib").import_module(pkg.__name__ + "._runtime")
    x = natural_frames(7, 24, 32, seed=3)
    frame_bytes = x[0].numel() * 4
    calls = []

    def fn(frames, first):
        calls.append((int(frames.shape[0]), int(first)))
        return frames + 0.25
    monkeypatch.setenv("VRGDG_STREAM_CHUNK_BYTES", str(2 * frame_bytes + 5))
    for stage in ("1", "0"):
        monkeypatch.setenv("VRGDG_STAGE_PAGEABLE", stage)
        calls.clear()
        out = rt.stream_frames(x, fn, 0, torch.device("cpu"), cuda_device)
        assert calls == [(2, 0), (2, 2), (2, 4), (1, 6)] and torch.equal(out, x + 0.25)
    monkeypatch.setenv("VRGDG_STREAM_CHUNK_BYTES", "0")             # cap off: the caller's chunk stands
    calls.clear()
    out = rt.stream_frames(x, fn, 0, torch.device("cpu"), cuda_device)
    assert calls == [(7, 0)] and torch.equal(out, x + 0.25)
    monkeypatch.delenv("VRGDG_STREAM_CHUNK_BYTES")
    monkeypatch.delenv("VRGDG_STAGE_PAGEABLE")
    lut = _lut33(pkg)
    node_out = pkg.VRGDG_LUTS().apply_lut(x, "B200 Vintage 33.cube", "auto", 10.0)[0]
    dev_out = pkg.VRGDG_LUTS._apply_cube_lut(x.to(cuda_device), lut["lut"], lut["domain_min"], lut["domain_max"]).cpu()
    assert node_out.device.type == "cpu" and node_out.is_pinned() and torch.equal(node_out, dev_out)


# ------------------------------------------------------------------------------------------------------
# graph-reachable nodes added this round
# ------------------------------------------------------------------------------------------------------
def test_postchain_node_equals_the_four_stock_nodes(pkg, cuda_device, oracle):
    node = pkg.NODE_CLASS_MAPPINGS["VRGDG_B200_PostChain"]()
    x = natural_frames(3, 72, 96, seed=8)
    ref_img = natural_frames(1, 50, 60, seed=9) * 0.8
    # without grain the node is deterministic: compare with the oracle composition colour match -> LUT -> unsharp
    out = node.apply_chain(x, 0.0, 0.5, 0.7, "B200 Vintage 33.cube", 6.0, "unsharp", 0.5, False, 2, reference_image=ref_img)[0]
    want = oracle.unsharp_numpy(oracle.apply_lut(oracle.color_match(x, ref_img, 0.7, 1), oracle.parse_cube(os.path.join(LUTS, "B200 Vintage 33.cube")), 6.0), 0.5)
    assert out.device.type == "cpu" and maxdiff(out, want) <= TOL
    # and the same as the stock node classes applied one after the other
    b = pkg.ColorMatchToReference().match_color(x, ref_img, 0.7, 1)[0]
    c = pkg.VRGDG_LUTS().apply_lut(b, "B200 Vintage 33.cube", "auto", 6.0)[0]
    d = pkg.FastUnsharpSharpen().apply_unsharp(c, 0.5, False)[0]
    assert maxdiff(out, d) <= 2e-6
    # with grain: reproducible under torch.manual_seed like FastFilmGrain, independent of the upload chunking
    torch.manual_seed(5)
    g1 = node.apply_chain(x, 0.04, 0.5, 1.0, "none", 10.0, "laplacian", 0.3, True, 1)[0]
    torch.manual_seed(5)
    g2 = node.apply_chain(x, 0.04, 0.5, 1.0, "none", 10.0, "laplacian", 0.3, True, 0)[0]
    assert torch.equal(g1, g2) and not torch.equal(g1, x)
    # nothing enabled -> the input itself
    assert node.apply_chain(x, 0.0, 0.5, 1.0, "none", 10.0, "none", 0.5, False, 8)[0] is x


def test_enhance_frames_node_equals_effects_batch(pkg, cuda_device):
    g = load_golden("effects")
    vt = __import__("importlib").import_module(pkg.__name__ + ".video_tools")
    node = pkg.NODE_CLASS_MAPPINGS["VRGDG_B200_EnhanceFrames"]()
    xe = t(g["xe"])
    out = node.enhance(xe, 0.8, 0.04, 0.5, 42, 7, True)[0]
    want = vt._apply_effects_batch(xe, {"sharpen_enabled": True, "sharpen_strength": 0.8, "grain_enabled": True, "grain_intensity": 0.04,
                                        "saturation_mix": 0.5, "seed": 42, "use_gpu": True}, 7)
    assert torch.equal(out, want)
    assert torch.equal(node.enhance(xe, 0.8, 0.0, 0.5, 42, 7, False)[0], t(g["sharp_only"]))        # the reference's unsharp output (use_gpu False)
    assert torch.equal(node.enhance(xe, 0.8, 0.04, 0.5, 42, 7, False)[0], vt._apply_effects_batch(xe, {"sharpen_enabled": True, "sharpen_strength": 0.8,
                       "grain_enabled": True, "grain_intensity": 0.04, "saturation_mix": 0.5, "seed": 42, "use_gpu": False}, 7))


def test_restore_original_node_vs_reference_outputs(pkg, cuda_device):
    import json
    from helpers import GOLDEN
    g = load_golden("restore_node")
    with open(os.path.join(GOLDEN, "restore_node_cases.json")) as fh:
        cases = json.load(fh)
    node = pkg.NODE_CLASS_MAPPINGS["VRGDGVideoEnhanceRestoreOriginal"]()
    for ci, (fit, method, strength) in enumerate(cases):
        ctx = {"original_frames": t(g["originals"]), "source_height": 30, "source_width": 40, "frame_count": 5, "fit_mode": fit, "fps": 24.0}
        frames, n, w, h, fps = node.restore(t(g["ltx"]), ctx, method, strength)
        assert (n, w, h, fps) == (5, 40, 30, 24.0) and frames.device.type == "cpu"
        tol = 0.0 if method in ("Nearest", "Area") else 2e-6                      # bilinear / bicubic: ATen's own kernels differ by that much
        assert maxdiff(frames, t(g["case%d" % ci])) <= tol, (fit, method)
    with pytest.raises(ValueError):
        node.restore(t(g["ltx"])[:1], dict(ctx, frame_count=20), "Bilinear", 1.0)
    assert pkg.NODE_CLASS_MAPPINGS["VRGDGStandaloneVideoEnhancer"]().return_output(None) == ("",)


# ------------------------------------------------------------------------------------------------------
# configs[4]: temporal 3-frame sharpen — a labelled extension (no reference operator exists; the NumPy oracle IS the specification)
# ------------------------------------------------------------------------------------------------------
def test_temporal_sharpen_vs_spec_oracle(pkg, cuda_device, oracle):
    x = natural_frames(6, 37, 53, seed=70)                       # odd sizes: scalar path
    x[3] = natural_frames(1, 37, 53, seed=71)[0]                 # a scene cut
    got = pkg.ops.temporal_sharpen(x.to(cuda_device), 0.7)
    assert torch.equal(got.cpu(), oracle.temporal_sharpen(x, 0.7))
    xv = natural_frames(5, 40, 64, seed=72)                      # vector path
    want = oracle.temporal_sharpen(xv, 1.3)
    assert torch.equal(pkg.ops.temporal_sharpen(xv.to(cuda_device), 1.3).cpu(), want)
    # a single frame and an empty clip
    assert torch.equal(pkg.ops.temporal_sharpen(xv[:1].to(cuda_device), 1.3).cpu(), oracle.temporal_sharpen(xv[:1], 1.3))
    assert pkg.ops.temporal_sharpen(xv[:0].to(cuda_device), 1.3).shape == (0, 40, 64, 3)
    # shards with halo frames == the whole clip (frame-sharded ranks exchange one frame per boundary, dist.exchange_halo_frames)
    d = xv.to(cuda_device)
    a = pkg.ops.temporal_sharpen(d[:2], 1.3, None, d[2])
    b = pkg.ops.temporal_sharpen(d[2:], 1.3, d[1], None)
    assert torch.equal(torch.cat([a, b]).cpu(), want)
    # 16-bit frames: fp32 arithmetic on the up-cast input, one rounding
    for dt in (torch.float16, torch.bfloat16):
        xh = xv.to(dt)
        assert torch.equal(pkg.ops.temporal_sharpen(xh.to(cuda_device), 1.3).cpu(), oracle.temporal_sharpen(xh.float(), 1.3).to(dt))
    # uint8 BGR wire format == decode -> spec -> encode
    u8 = (xv * 255).to(torch.uint8).flip(-1).contiguous()
    dec = oracle.frames_to_tensor(list(u8.numpy()))
    enc = np.stack(oracle.tensor_to_frames(oracle.temporal_sharpen(dec, 1.3)))
    assert np.array_equal(pkg.ops.temporal_sharpen(u8.to(cuda_device), 1.3).cpu().numpy(), enc)
    # the node chunks a host clip and carries the neighbours along
    node = pkg.NODE_CLASS_MAPPINGS["VRGDG_B200_TemporalSharpen"]()
    assert torch.equal(node.sharpen(xv, 1.3, 2)[0], want) and torch.equal(node.sharpen(xv, 1.3, 0)[0], want)
    with pytest.raises(ValueError):
        pkg.ops.temporal_sharpen(d, 1.3, d[0, :10], None)


def test_temporal_sharpen_full_size_config5_shape(pkg, cuda_device, oracle):
    """configs[4] frame size (1080p fp32) at a reduced clip length against the spec oracle, plus size-independent properties"""
    x = natural_frames(8, 1080, 1920, seed=73)
    got = pkg.ops.temporal_sharpen(x.to(cuda_device), 0.5)
    assert torch.equal(got.cpu(), oracle.temporal_sharpen(x, 0.5))
    still = x[:1].repeat(6, 1, 1, 1).to(cuda_device)               # a still clip is a fixed point up to the rounding of (3x)/3
    assert maxdiff(pkg.ops.temporal_sharpen(still, 2.0), still) <= 3e-7
    assert torch.equal(pkg.ops.temporal_sharpen(x.to(cuda_device), 0.0).cpu(), x)      # strength 0 = identity


# ------------------------------------------------------------------------------------------------------
# histogram / CDF colour transfer — a labelled extension (the reference has no histogram; the NumPy oracle IS the specification)
# ------------------------------------------------------------------------------------------------------
def test_histogram_colormatch_vs_spec_oracle(pkg, cuda_device, oracle):
    x = natural_frames(3, 72, 93, seed=80)
    x[1] = (x[1] * 0.6 + 0.3).clamp(0, 1)
    x[0, :2, :5] = torch.tensor([0.0, 1.0, 0.5])              # exact bin edges / extremes
    ref = natural_frames(1, 50, 61, seed=81) * 0.8 + 0.05
    xd, rd = x.to(cuda_device), ref.to(cuda_device)
    counts = pkg.ops.hist_counts(xd)
    want_counts = oracle.hist_counts(x)
    assert torch.equal(counts.cpu().long(), want_counts)       # integers: exact
    assert int(counts[0].sum()) == 3 * 72 * 93
    # row shards add up to the whole image (what the ranks exchange: dist.reference_histogram_distributed)
    parts = pkg.ops.hist_counts(rd, 0, 20) + pkg.ops.hist_counts(rd, 20, 30)
    assert torch.equal(parts, pkg.ops.hist_counts(rd))
    tables = pkg.ops.histmatch_tables(counts, pkg.ops.hist_counts(rd))
    T = oracle.histmatch_tables(want_counts, oracle.hist_counts(ref))
    assert torch.equal(tables[..., 0].cpu(), T[..., :256])     # fp64 edge arithmetic with one rounding per op: bit-identical
    assert torch.equal(tables[..., 1].cpu(), T[..., 1:] - T[..., :256])
    assert bool((tables[..., 1] >= 0).all())                   # monotone
    for strength in (1.0, 0.35):
        got = pkg.ops.histmatch_apply(xd, tables, strength, 1.0 - strength)
        assert torch.equal(got.cpu(), oracle.hist_match(x, ref, strength))
    # the node, host tensors in / out, chunked
    node = pkg.NODE_CLASS_MAPPINGS["VRGDG_B200_HistogramColorMatch"]()
    assert torch.equal(node.match_histogram(x, ref, 0.35, 2)[0], oracle.hist_match(x, ref, 0.35))
    # 16-bit and uint8 BGR frames: fp32 arithmetic on the decoded values, one rounding / the truncating encode
    for dt in (torch.float16, torch.bfloat16):
        xh, rh = x.to(dt), ref.to(dt)
        th = pkg.ops.histmatch_tables(pkg.ops.hist_counts(xh.to(cuda_device)), pkg.ops.hist_counts(rh.to(cuda_device)))
        assert torch.equal(pkg.ops.histmatch_apply(xh.to(cuda_device), th, 1.0, 0.0).cpu(), oracle.hist_match(xh.float(), rh.float(), 1.0).to(dt))
    u8 = (x * 255).to(torch.uint8).flip(-1).contiguous()
    r8 = (ref * 255).to(torch.uint8).flip(-1).contiguous()
    dec, rdec = oracle.frames_to_tensor(list(u8.numpy())), oracle.frames_to_tensor(list(r8.numpy()))
    t8 = pkg.ops.histmatch_tables(pkg.ops.hist_counts(u8.to(cuda_device)), pkg.ops.hist_counts(r8.to(cuda_device)))
    enc = np.stack(oracle.tensor_to_frames(oracle.hist_match(dec, rdec, 1.0)))
    assert np.array_equal(pkg.ops.histmatch_apply(u8.to(cuda_device), t8, 1.0, 0.0).cpu().numpy(), enc)
    # empty reference rows -> identity tables; one reference per frame
    ident = pkg.ops.histmatch_tables(counts, torch.zeros_like(counts[:1]))
    assert maxdiff(pkg.ops.histmatch_apply(xd, ident, 1.0, 0.0), xd.clamp(0, 1)) <= 1e-6
    per_frame = pkg.ops.histmatch_tables(counts, counts)       # every frame matched to itself
    assert per_frame.shape == (3, 3, 256, 2)


def test_histogram_colormatch_full_size_properties(pkg, cuda_device):
    """configs[2] frame size: one 4K frame matched to a 4K reference lands on the reference's CDF (size-independent property)"""
    x = natural_frames(1, 2160, 3840, seed=82, device=cuda_device)
    ref = (natural_frames(1, 2160, 3840, seed=83, device=cuda_device) * 0.7 + 0.2).clamp(0, 1)
    cx, cr = pkg.ops.hist_counts(x), pkg.ops.hist_counts(ref)
    assert int(cx.sum()) == 3 * 2160 * 3840
    out = pkg.ops.histmatch_apply(x, pkg.ops.histmatch_tables(cx, cr), 1.0, 0.0)
    co = pkg.ops.hist_counts(out)
    n = 2160.0 * 3840.0
    cdf_o, cdf_r = co.double().cumsum(-1) / n, cr.double().cumsum(-1) / n
    assert float((cdf_o - cdf_r).abs().max()) < 0.02           # matched CDF follows the reference CDF (bin-quantisation slack)
    # monotone: a brighter input never gets darker within a channel
    idx = torch.argsort(x[0, :, :, 1].flatten()[:200000])
    assert bool((out[0, :, :, 1].flatten()[:200000][idx].diff() >= -1e-6).all())
    same = pkg.ops.histmatch_apply(x, pkg.ops.histmatch_tables(cx, cx), 1.0, 0.0)
    assert maxdiff(same, x) <= 1e-6                            # a frame matched to itself is the identity (plateau rule of the inverse CDF)
