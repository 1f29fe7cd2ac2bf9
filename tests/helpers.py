"""Shared test helpers: synthetic frames and golden loading."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LUTS = os.path.join(ROOT, "comfyui-vrgamedevgirl_b200", "LUTS")


def white_frames(B, H, W, seed=1, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, H, W, 3, generator=g).to(dtype)


def natural_frames(B, H, W, seed=0, dtype=torch.float32, device="cpu"):
    """Spatially coherent frames: 4 octaves of bilinearly upsampled uniform noise + 2% white noise (dist. N)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    acc = torch.zeros(B, 3, H, W)
    amp, tot = 1.0, 0.0
    for o in range(4):
        gh, gw = 8 * 2 ** o + 1, 15 * 2 ** o + 1
        base = torch.rand(B, 3, gh, gw, generator=g)
        acc += amp * torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=True)
        tot += amp
        amp *= 0.5
    acc = acc / tot
    acc = acc + 0.02 * (torch.rand(B, 3, H, W, generator=g) - 0.5)
    off = torch.tensor([0.03, 0.0, -0.03]).view(1, 3, 1, 1)
    return (acc + off).clamp(0, 1).permute(0, 2, 3, 1).contiguous().to(dtype).to(device)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
