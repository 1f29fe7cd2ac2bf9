"""Host-side logic that mirrors the reference without needing a GPU: node API surface, .cube parsing, palette LUT,
file naming, batch-size table, and the 'no CUDA -> raise' contract."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, LUTS, load_golden, t, white_frames


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(GOLDEN, "reference_meta.json"), encoding="utf-8") as fh:
        return json.load(fh)


EXTRA_KEYS = {"VRGDG_B200_PostChain", "VRGDG_B200_EnhanceFrames", "VRGDG_B200_TemporalSharpen", "VRGDG_B200_HistogramColorMatch"}      # this package's own nodes (no reference counterpart)


def test_node_mappings_and_api_match_reference(pkg, meta):
    with open(os.path.join(GOLDEN, "reference_meta_r2.json"), encoding="utf-8") as fh:
        meta2 = json.load(fh)                                              # the two "next"-row reference nodes (make_golden_r2.py)
    api = dict(meta["api"], **meta2["api"])
    assert set(pkg.NODE_CLASS_MAPPINGS) == set(api) | EXTRA_KEYS
    for key, want in api.items():
        cls = pkg.NODE_CLASS_MAPPINGS[key]
        got = json.loads(json.dumps(cls.INPUT_TYPES()))
        if key == "VRGDG_LUTS":
            assert got["required"]["lut_name"][0] == sorted(got["required"]["lut_name"][0], key=str.lower)
            got["required"]["lut_name"] = ["<lut files>"]
        assert got == want["INPUT_TYPES"], key
        assert list(got["required"]) == list(want["INPUT_TYPES"]["required"]), key + " widget order"
        assert list(cls.RETURN_TYPES) == want["RETURN_TYPES"] and cls.FUNCTION == want["FUNCTION"] and cls.CATEGORY == want["CATEGORY"]
        assert list(getattr(cls, "RETURN_NAMES", ())) == want["RETURN_NAMES"]
        assert getattr(cls, "DESCRIPTION", None) == want["DESCRIPTION"]
        assert callable(getattr(cls, cls.FUNCTION))
        assert bool(getattr(cls, "OUTPUT_NODE", False)) == bool(want.get("OUTPUT_NODE", False))
    names = dict(meta["display_names"], **meta2["display_names"])
    assert {k: v for k, v in pkg.NODE_DISPLAY_NAME_MAPPINGS.items() if k not in EXTRA_KEYS} == names
    for key in EXTRA_KEYS:                                                   # ComfyUI's node contract for the extra keys
        cls = pkg.NODE_CLASS_MAPPINGS[key]
        it = cls.INPUT_TYPES()
        assert "images" in it["required"] and it["required"]["images"] == ("IMAGE",)
        assert cls.RETURN_TYPES == ("IMAGE",) and callable(getattr(cls, cls.FUNCTION)) and isinstance(cls.CATEGORY, str)
        assert key in pkg.NODE_DISPLAY_NAME_MAPPINGS
        import inspect
        params = list(inspect.signature(getattr(cls, cls.FUNCTION)).parameters)[1:]
        assert params == list(it["required"]) + list(it.get("optional", {})), key      # widgets are passed by name, in this order


def test_method_signatures_are_the_reference_ones(pkg):
    import inspect
    sig = lambda c, m: list(inspect.signature(getattr(c, m)).parameters)
    assert sig(pkg.FastFilmGrain, "apply_grain") == ["self", "images", "grain_intensity", "saturation_mix", "batch_size"]
    assert sig(pkg.ColorMatchToReference, "match_color") == ["self", "images", "reference_image", "match_strength", "batch_size"]
    for c, m in ((pkg.FastUnsharpSharpen, "apply_unsharp"), (pkg.FastLaplacianSharpen, "apply_laplacian"), (pkg.FastSobelSharpen, "apply_sobel")):
        assert sig(c, m) == ["self", "images", "strength", "use_gpu"]
    assert sig(pkg.VRGDG_LUTS, "apply_lut") == ["self", "image", "lut_name", "device", "strength"]
    assert sig(pkg.VRGDG_MakeLUT, "create_and_apply") == ["self", "image", "colors", "name_suffix", "lut_size", "device", "strength"]
    # saved workflows store positional widget values, e.g. FastFilmGrain [0.01, 0.5, 4], ColorMatchToReference [1, 4]
    assert list(pkg.FastFilmGrain.INPUT_TYPES()["required"])[1:] == ["grain_intensity", "saturation_mix", "batch_size"]


def test_cube_parser_matches_oracle_and_rejects_bad_files(pkg, oracle, tmp_path):
    for fname in sorted(os.listdir(LUTS)) + [os.path.join(GOLDEN, "domain_5.cube")]:
        path = fname if os.path.isabs(fname) else os.path.join(LUTS, fname)
        if not path.endswith(".cube"):
            continue
        a, b = pkg.VRGDG_LUTS._parse_cube_file(path), oracle.parse_cube(path)
        assert a["size"] == b["size"] and torch.equal(a["lut"], b["lut"])
        assert torch.equal(a["domain_min"], b["domain_min"]) and torch.equal(a["domain_max"], b["domain_max"])
        assert a["lut"].dtype == torch.float32 and tuple(a["lut"].shape) == (a["size"],) * 3 + (3,)

    def write(name, text):
        p = tmp_path / name
        p.write_text(text)
        return str(p)
    body = "".join("0.1 0.2 0.3\n" for _ in range(8))
    ok = pkg.VRGDG_LUTS._parse_cube_file(write("ok.cube", "# c\nTITLE \"x\"\nLUT_3D_SIZE 2\nDOMAIN_MIN 0 0 0\nDOMAIN_MAX 1 1 1\nFOO 1\nLUT_3D_INPUT_RANGE 0.0 1.0 2.0 3.0\n\n" + body))
    assert ok["size"] == 2 and ok["lut"][1, 0, 1].tolist() == pytest.approx([0.1, 0.2, 0.3])
    # red fastest: value index 1 is [b=0][g=0][r=1]
    seq = "".join("%d 0 0\n" % i for i in range(8))
    lut = pkg.VRGDG_LUTS._parse_cube_file(write("seq.cube", "LUT_3D_SIZE 2\n" + seq))["lut"]
    assert lut[0, 0, 1, 0] == 1 and lut[0, 1, 0, 0] == 2 and lut[1, 0, 0, 0] == 4
    for name, text, exc in (("a.cube", "LUT_1D_SIZE 4\n", ValueError), ("b.cube", body, ValueError), ("c.cube", "LUT_3D_SIZE 2\n0.1 0.2 0.3\n", ValueError),
                            ("d.cube", "LUT_3D_SIZE 2 3\n" + body, ValueError), ("e.cube", "LUT_3D_SIZE 2\nDOMAIN_MIN 0 0\n" + body, ValueError)):
        with pytest.raises(exc):
            pkg.VRGDG_LUTS._parse_cube_file(write(name, text))
        with pytest.raises(exc):
            oracle.parse_cube(write(name, text))


def test_palette_lut_and_cube_writer_round_trip(pkg, tmp_path):
    ln = __import__("importlib").import_module("comfyui-vrgamedevgirl_b200.lut_nodes")
    g = load_golden("palette")
    assert torch.equal(ln._build_palette_lut("#0b1d51, #1f6aa5, #f3d27a", 9), t(g["three"]))
    assert torch.equal(ln._build_palette_lut("teal", 8), t(g["one"]))
    assert torch.equal(ln._build_palette_lut("black, #f80, white, pink", 11), t(g["names"]))
    with pytest.raises(ValueError):
        ln._build_palette_lut("#12345", 8)
    with pytest.raises(ValueError):
        ln._build_palette_lut(" , ", 8)
    path = str(tmp_path / "sub" / "p.cube")
    ln._write_cube_file(t(g["three"]), path)
    head = open(path).read().splitlines()[:4]
    assert head == ['TITLE "p.cube"', "LUT_3D_SIZE 9", "DOMAIN_MIN 0.0 0.0 0.0", "DOMAIN_MAX 1.0 1.0 1.0"]
    back = ln.VRGDG_LUTS._parse_cube_file(path)
    assert back["size"] == 9 and float((back["lut"] - t(g["three"])).abs().max()) <= 5.1e-7      # %.6f text
    assert ln._sanitize_filename_part(" #0B1d51 ") == "0b1d51" and ln._sanitize_filename_part("") == "custom"
    assert ln._sanitize_filename_part("My  Fancy--Name!") == "my_fancy_name"


def test_lut_listing_cache_and_is_changed(pkg, monkeypatch, tmp_path):
    ln = __import__("importlib").import_module("comfyui-vrgamedevgirl_b200.lut_nodes")
    names = ln._list_lut_files()
    assert "B200 Vintage 33.cube" in names and names == sorted(names, key=str.lower)
    d1 = pkg.VRGDG_LUTS._load_lut("B200 Vintage 33.cube")
    assert pkg.VRGDG_LUTS._load_lut("B200 Vintage 33.cube") is d1 and len(pkg.VRGDG_LUTS._LUT_CACHE) == 1
    key = pkg.VRGDG_LUTS.IS_CHANGED(None, "B200 Vintage 33.cube", "auto", 10.0)
    assert key.endswith("|auto|10.0") and "B200 Vintage 33.cube" in key
    assert pkg.VRGDG_LUTS.IS_CHANGED(None, "No LUT files found", "cpu", 1.0) == "missing|cpu|1.0"
    assert "|missing|nope.cube|" in pkg.VRGDG_LUTS.IS_CHANGED(None, "nope.cube", "cpu", 1.0)
    monkeypatch.setattr(ln, "LUTS_DIR", str(tmp_path / "none"))
    assert ln._list_lut_files() == ["No LUT files found"] and pkg.VRGDG_LUTS._get_luts_folder_state() == "missing"
    with pytest.raises(ValueError):
        pkg.VRGDG_LUTS._load_lut("No LUT files found")
    with pytest.raises(FileNotFoundError):
        pkg.VRGDG_LUTS._load_lut("x.cube")
    assert ln._next_available_lut_path("a_b").endswith("a_b.cube")
    open(os.path.join(str(tmp_path / "none"), "a_b.cube"), "w").close()
    assert ln._next_available_lut_path("a_b").endswith("a_b_2.cube")


def test_auto_batch_size_table(pkg, meta):
    vt = __import__("importlib").import_module("comfyui-vrgamedevgirl_b200.video_tools")
    for k, v in meta["auto_batch"].items():
        w, h = (int(s) for s in k.split("x"))
        assert vt._auto_batch_size(w, h) == v


def test_nodes_raise_without_cuda_instead_of_falling_back(pkg):
    if torch.cuda.is_available():
        pytest.skip("this contract is about machines without a GPU")
    x = white_frames(1, 8, 8)
    with pytest.raises(RuntimeError):
        pkg.FastFilmGrain().apply_grain(x, 0.04, 0.5, 4)
    with pytest.raises(RuntimeError):
        pkg.FastUnsharpSharpen().apply_unsharp(x, 0.5, False)
    with pytest.raises(RuntimeError):
        pkg.ColorMatchToReference().match_color(x, x, 1.0, 1)
    for dev in ("auto", "cuda", "cpu"):
        with pytest.raises(RuntimeError):
            pkg.VRGDG_LUTS().apply_lut(x, "B200 Vintage 33.cube", dev, 10.0)
    with pytest.raises(RuntimeError):
        pkg.ops.grain(x, 0.04, 0.5, 0.5, seed=1)


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: no file of the package may reference it"""
    from conftest import PKG_NAME, ROOT
    for dirpath, _, files in os.walk(os.path.join(ROOT, PKG_NAME)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "vrgdg_oracle" not in text and "ref_harness" not in text and "/root/reference" not in text, f


def test_u8_division_identity():
    """Elem<uint8_t>::ld computes v/255.0f as q' = fma(fma(-255, q, v), r, q), q = v*r, r = fl(1/255): must equal the IEEE quotient
    (what numpy's astype(float32)/255.0 produces) for all 256 byte values.  Exact rational arithmetic emulates the FMAs."""
    from fractions import Fraction

    def rn32(fr):
        c = np.float32(float(fr))
        cands = [np.nextafter(c, np.float32(-np.inf), dtype=np.float32), c, np.nextafter(c, np.float32(np.inf), dtype=np.float32)]
        return min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(np.float32(v).view(np.uint32)) & 1))

    r = np.float32(1.0) / np.float32(255.0)
    for v in range(256):
        q = rn32(Fraction(v) * Fraction(float(r)))
        res = rn32(Fraction(v) - Fraction(255) * Fraction(float(q)))
        q2 = rn32(Fraction(float(res)) * Fraction(float(r)) + Fraction(float(q)))
        assert q2 == np.float32(v) / np.float32(255.0), v


def test_adjust_settings_normalisation_and_descriptor(pkg, oracle):
    vt = __import__("importlib").import_module("comfyui-vrgamedevgirl_b200.video_tools")
    for st in ({}, {"enabled": False}, {"sharpen": -3, "fade": 1e9, "tint": "x"}, {"clarity": 0.05}, None, "junk"):
        assert vt._normalize_adjust_settings(st) == oracle.normalize_adjust_settings(st)
    d = vt._adjust_desc({"temperature": 40, "tint": -9, "exposure": 50, "clarity": 0.05, "sharpen": 0.2, "fade": 10, "vignette": 0}, 1080, 1920)
    assert d.enabled == 1 and d.clarity_on == 0 and d.sharpen_on == 1 and d.blur_kernel == 9 and d.fade_on == 1 and d.vignette_on == 0
    assert d.exposure == np.float32(2.0 ** 0.5) and d.offset_rgb[0] == np.float32(40 / 400.0 - (-9) / 900.0) and d.fade_mul == np.float32(1.0 - 0.1 * 0.35)
    assert vt._adjust_desc({}, 5, 8).blur_kernel == 5 and vt._adjust_desc({}, 2, 2).blur_kernel == 1 and vt._adjust_desc({}, 4, 100).blur_kernel == 3
    lib = pkg._native.load_library()
    import ctypes
    assert lib.vrgdg_adjust_scratch_bytes(2, 10, 20, ctypes.byref(d)) == 2 * 10 * 20 * 3 * 4
    assert lib.vrgdg_adjust_scratch_bytes(2, 10, 20, ctypes.byref(vt._adjust_desc({}, 10, 20))) == 0


def test_resize_plan_matches_the_reference_shapes(pkg, oracle):
    """video_enhance._resize_plan / _output_size (host arithmetic of _resize_batch :66-85): for random sizes the planned output
    shape, content rectangle and offsets equal what the reference's interpolate + slice / pad produce (oracle, nearest, tiny)."""
    import importlib
    ve = importlib.import_module("comfyui-vrgamedevgirl_b200.video_enhance")
    rng = np.random.default_rng(5)
    for _ in range(60):
        sw, sh, tw, th = (int(v) for v in rng.integers(1, 40, size=4))
        x = torch.ones(1, sh, sw, 3)
        for fit in ("Stretch to dimensions", "Crop to fill", "Fit with letterbox (preserve all)"):
            ref = oracle.resize_batch(x, tw, th, fit, "Nearest")
            res, off = ve._resize_plan(sw, sh, tw, th, fit)
            ow, oh = ve._output_size(res, off, tw, th, fit)
            assert (oh, ow) == tuple(ref.shape[1:3]), (sw, sh, tw, th, fit)
            inside = torch.zeros(oh, ow)
            inside[max(off[1], 0):max(off[1], 0) + min(res[1], oh), max(off[0], 0):max(off[0], 0) + min(res[0], ow)] = 1
            assert torch.equal(inside, ref[0, :, :, 0]), (sw, sh, tw, th, fit)
    assert ve._interpolation("Area") == "area" and ve._interpolation("nope") == "bicubic"
    with pytest.raises(ValueError):
        ve._resize_batch(torch.zeros(0, 4, 4, 3), 8, 8, "Stretch to dimensions", "Nearest")


def test_lanczos4_host_tables_match_the_oracle(pkg, oracle):
    """vrgdg_lanczos4_tables is host code (no device work): OpenCV's 8-tap fixed-point tables, incl. the two size pairs where
    rounding x + 3 in fp32 (as OpenCV does) rather than fp64 changes a weight (317 -> 2252 index 17, 500 -> 378 index 0)."""
    import ctypes
    lib = pkg._native.load_library()
    for s_, d_ in ((317, 2252), (500, 378), (53, 106), (7, 3), (1, 5), (720, 1080), (1920, 1280), (64, 64), (1080, 2160)):
        ofs, coef = np.empty(d_, np.int32), np.empty((d_, 8), np.int16)
        assert lib.vrgdg_lanczos4_tables(s_, d_, ofs.ctypes.data_as(ctypes.c_void_p), coef.ctypes.data_as(ctypes.c_void_p)) == 0
        o2, c2 = oracle.lanczos4_tables(s_, d_)
        assert np.array_equal(ofs, o2) and np.array_equal(coef, c2), (s_, d_)
    ofs, coef = np.empty(64, np.int32), np.empty((64, 8), np.int16)
    lib.vrgdg_lanczos4_tables(64, 64, ofs.ctypes.data_as(ctypes.c_void_p), coef.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(ofs, np.arange(64)) and np.array_equal(coef, np.tile(np.array([0, 0, 0, 2048, 0, 0, 0, 0], np.int16), (64, 1)))
    assert lib.vrgdg_lanczos4_tables(0, 4, ofs.ctypes.data_as(ctypes.c_void_p), coef.ctypes.data_as(ctypes.c_void_p)) == -1
    assert lib.vrgdg_lanczos4_scratch_bytes(2, 10, 7) == 2 * 10 * 7 * 3 * 4


def test_random_cube_files_and_lanczos_tables_fuzz(pkg, oracle, tmp_path):
    """Seeded fuzz of the two host-side table builders against the oracle: .cube text with random whitespace / comments / keyword
    lines / exponent notation / non-unit domains (product parser == oracle parser == the values written), and Lanczos4 tables for
    random size pairs (the library's host code == the oracle's restatement of OpenCV)."""
    import ctypes
    rng = np.random.default_rng(99)
    for n in range(12):
        S = int(rng.integers(2, 7))
        vals = rng.random((S, S, S, 3)).astype(np.float32) * float(rng.choice([1.0, 2.0, 0.5])) - float(rng.choice([0.0, 0.25]))
        lines = ["# fuzz %d" % n, 'TITLE "t %d"' % n, ("LUT_3D_SIZE %d" % S) if n % 2 else ("LUT_3D_SIZE\t%d  " % S)]
        dom = n % 3 == 0
        if dom:
            lines += ["DOMAIN_MIN -0.25 0 0.5", "DOMAIN_MAX 1.5 1 2e0"]
        if n % 4 == 1:
            lines += ["LUT_3D_INPUT_RANGE 0.0 1.0 2.0", "SOMETHING else entirely here 1 2"]     # not 3 tokens -> skipped like the reference
        for i, v in enumerate(vals.reshape(-1, 3)):
            fmt = ("%.6f %.6f %.6f", "%.9g\t%.9g   %.9g", "  %e %e %e  ")[(n + i) % 3]
            lines.append(fmt % tuple(float(c) for c in v))
            if i % 11 == 5:
                lines += ["", "# comment in the table"]
        path = tmp_path / ("fuzz %d.cube" % n)
        path.write_text("\n".join(lines) + "\n")
        a, b = pkg.VRGDG_LUTS._parse_cube_file(str(path)), oracle.parse_cube(str(path))
        assert a["size"] == b["size"] == S and torch.equal(a["lut"], b["lut"])
        assert torch.equal(a["domain_min"], b["domain_min"]) and torch.equal(a["domain_max"], b["domain_max"])
        assert a["domain_max"].tolist() == ([1.5, 1.0, 2.0] if dom else [1.0, 1.0, 1.0])
        assert np.abs(a["lut"].numpy() - vals).max() <= 1e-6                 # %.6f rounding at most; red fastest order preserved
    bad = tmp_path / "three_token_keyword.cube"                              # a 3-token non-numeric line is read as data by the reference
    bad.write_text("LUT_3D_SIZE 2\nLUT_3D_INPUT_RANGE 0.0 1.0\n" + "0 0 0\n" * 8)
    for parse in (pkg.VRGDG_LUTS._parse_cube_file, oracle.parse_cube):
        with pytest.raises(ValueError):
            parse(str(bad))
    lib = pkg._native.load_library()
    for _ in range(25):
        s_, d_ = int(rng.integers(1, 1500)), int(rng.integers(1, 1500))
        ofs, coef = np.empty(d_, np.int32), np.empty((d_, 8), np.int16)
        assert lib.vrgdg_lanczos4_tables(s_, d_, ofs.ctypes.data_as(ctypes.c_void_p), coef.ctypes.data_as(ctypes.c_void_p)) == 0
        o2, c2 = oracle.lanczos4_tables(s_, d_)
        assert np.array_equal(ofs, o2) and np.array_equal(coef, c2), (s_, d_)
        assert int(np.abs(coef.astype(np.int32).sum(axis=1) - 2048).max()) <= 3     # weights sum to 1 in fixed point, up to rounding


def test_pipeline_chunk_rule(pkg, monkeypatch):
    """host batches are cut into pipeline chunks by BYTES whatever the node's batch widget says (results never depend on the cut);
    the rule is a pure function of (caller's chunk, frame size, cap)"""
    rt = __import__("importlib").import_module(pkg.__name__ + "._runtime")
    frame_4k = 2160 * 3840 * 3 * 4
    monkeypatch.delenv("VRGDG_STREAM_CHUNK_BYTES", raising=False)
    assert rt.pipeline_chunk(16, frame_4k) == 2                      # 256 MiB / 99.5 MB
    assert rt.pipeline_chunk(1, frame_4k) == 1                       # the caller's smaller chunk stands
    assert rt.pipeline_chunk(500, 1920 * 1080 * 3 * 2) == 21         # 1080p fp16
    assert rt.pipeline_chunk(8, 1 << 30) == 1                        # a frame larger than the cap still moves, one at a time
    assert rt.pipeline_chunk(8, frame_4k, cap=0) == 8                # cap off
    monkeypatch.setenv("VRGDG_STREAM_CHUNK_BYTES", str(4 * frame_4k))
    assert rt.pipeline_chunk(16, frame_4k) == 4
    monkeypatch.setenv("VRGDG_STREAM_CHUNK_BYTES", "not a number")   # unreadable value: the default
    assert rt.pipeline_chunk(16, frame_4k) == 2
