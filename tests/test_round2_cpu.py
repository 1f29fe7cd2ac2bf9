"""Round-2 CPU tests: the oracle and the host build of the kernels' arithmetic against the reference's outputs on 64^3 / 65^3
tables (28 of the 40 LUT files the reference ships have these sizes), golden tests/golden/lut_big.npz (make_golden_r2.py)."""
import ctypes
import hashlib
import os

import numpy as np
import torch

from helpers import load_golden, t, write_big_cube


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_big_lut_tables_parse_identically_and_oracle_is_bit_exact(oracle, pkg, tmp_path):
    g = load_golden("lut_big")
    for size in (64, 65):
        path = write_big_cube(str(tmp_path / ("big_%d.cube" % size)), size)
        d = oracle.parse_cube(path)
        ours = pkg.VRGDG_LUTS._parse_cube_file(path)             # the product's parser (host code, no GPU needed)
        assert d["size"] == size and ours["size"] == size
        assert np.array_equal(_sha(d["lut"].numpy()), g["sha_%d" % size]), "generated table differs from the one the golden was made from"
        assert torch.equal(ours["lut"], d["lut"]) and torch.equal(ours["domain_min"], d["domain_min"]) and torch.equal(ours["domain_max"], d["domain_max"])
        assert torch.equal(oracle.apply_lut(t(g["x"]), d, 10.0), t(g["s%d__s10" % size]))
        assert torch.equal(oracle.apply_lut(t(g["x"]), d, 3.5), t(g["s%d__s3p5" % size]))
        assert torch.equal(oracle.apply_lut(t(g["xn"]), d, 10.0), t(g["s%d__nat" % size]))
        assert torch.equal(oracle.apply_lut(t(g["x"]).half(), d, 10.0), t(g["s%d__fp16" % size]))


def test_big_lut_kernel_arithmetic_bit_exact_on_host(hostcheck, oracle, tmp_path):
    vp, i64, ci, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    hostcheck.hc_lut3d.argtypes = [vp, vp, i64, vp, ci, vp, vp, f32, f32, ci]
    P = lambda a: a.ctypes.data_as(vp)
    g = load_golden("lut_big")
    x = np.ascontiguousarray(g["x"].reshape(-1, 3))
    o = np.zeros_like(x)
    for size in (64, 65):
        d = oracle.parse_cube(write_big_cube(str(tmp_path / ("big_%d.cube" % size)), size))
        lut = np.ascontiguousarray(d["lut"].numpy())
        dmin = d["domain_min"].numpy().copy()
        span = torch.clamp(d["domain_max"] - d["domain_min"], min=1e-6).numpy().copy()
        hostcheck.hc_lut3d(P(x), P(o), x.shape[0], P(lut), size, P(dmin), P(span), 1.0, 0.0, 1)
        assert np.array_equal(o, g["s%d__s10" % size].reshape(-1, 3))
        hostcheck.hc_lut3d(P(x), P(o), x.shape[0], P(lut), size, P(dmin), P(span), 0.35, 1.0 - 0.35, 3)
        assert np.array_equal(o, g["s%d__s3p5" % size].reshape(-1, 3))
