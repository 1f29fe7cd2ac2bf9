"""Round-2 additions to the golden vectors, produced by EXECUTING THE REFERENCE'S OWN SOURCE (oracle/ref_harness.py).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_r2.py
  lut_big.npz : VRGDG_LUTS._parse_cube_file + _apply_cube_lut + the strength blend of apply_lut on generated 64^3 / 65^3 tables
                (tests/helpers.py::write_big_cube; the 65^3 file carries integer DOMAIN_MIN/MAX lines)
"""
import hashlib
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_harness as RH  # noqa: E402
from helpers import natural_frames, white_frames, write_big_cube  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_num_threads(4)


def main():
    assert RH.available(), "reference tree not found"
    iv = RH.load_iv_adjustments()
    out = {}
    xw = white_frames(1, 48, 64, seed=31)
    xw[0, 0, 0] = torch.tensor([1.0, 0.0, 1.0])
    xw[0, 0, 1] = torch.tensor([0.0, 0.0, 0.0])
    xw[0, 0, 2] = torch.tensor([1.0, 1.0, 1.0])
    xw[0, 0, 3] = torch.tensor([0.5, 0.25, 0.75])                     # exact grid nodes of the 65^3 table
    xw[0, 0, 4] = torch.tensor([1.0 / 63.0, 62.0 / 63.0, 0.5])        # grid nodes / mid-cell of the 64^3 table
    xn = natural_frames(1, 48, 64, seed=32)
    out["x"], out["xn"] = xw, xn
    with tempfile.TemporaryDirectory() as tmp:
        iv.LUTS_DIR = tmp
        node = iv.VRGDG_LUTS()
        for size in (64, 65):
            name = "big_%d.cube" % size
            write_big_cube(os.path.join(tmp, name), size)
            data = iv.VRGDG_LUTS._parse_cube_file(os.path.join(tmp, name))
            assert data["size"] == size
            out["sha_%d" % size] = np.frombuffer(hashlib.sha256(data["lut"].contiguous().numpy().tobytes()).digest(), dtype=np.uint8)
            out["s%d__s10" % size] = node.apply_lut(xw, name, "cpu", 10.0)[0]
            out["s%d__s3p5" % size] = node.apply_lut(xw, name, "cpu", 3.5)[0]
            out["s%d__nat" % size] = node.apply_lut(xn, name, "cpu", 10.0)[0]
            out["s%d__fp16" % size] = node.apply_lut(xw.half(), name, "cpu", 10.0)[0]
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "lut_big.npz"), **arrays)
    print("lut_big", {k: tuple(v.shape) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
