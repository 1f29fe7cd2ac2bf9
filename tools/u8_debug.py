"""Debug helper (GPU box): uint8 stencil through the generic loader, then through TMA."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("comfyui-vrgamedevgirl_b200")
from helpers import load_golden, t
nv = pkg._native
g = load_golden("u8chain")
dev = torch.device("cuda", 0)
x = t(g["bgr_in"]).to(dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "generic"
if mode == "generic":
    os.environ["VRGDG_NO_TMA"] = "1"
u = pkg.ops.stencil3x3(x, nv.STENCIL_BOX_UNSHARP, 0.5, nv.BORDER_REPLICATE)
torch.cuda.synchronize()
print(mode, nv.last_tile_path(), "equal to golden:", torch.equal(u.cpu(), t(g["unsharp_only"])), "maxdiff", int((u.cpu().int() - t(g["unsharp_only"]).int()).abs().max()))
