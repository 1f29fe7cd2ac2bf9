"""Fused post-processing chain: grain -> colour match -> 3D LUT -> 3x3 stencil -> post grain in ONE pass over HBM.

Semantics = the reference nodes applied one after another (FastFilmGrain nodes.py:41-66 ->
ColorMatchToReference :91-124 -> VRGDG_LUTS VRGDG_IV_Adjustments.py:345-361 -> FastUnsharpSharpen nodes.py:156-209),
or the standalone enhancer's unsharp -> seeded grain (_apply_effects_batch,
VRGDG_StandaloneVideoEnhancerNodes.py:278-294) via `post_grain`.  Any subset of stages may be enabled.
"""
import ctypes

import torch

from . import _native as nv
from . import ops
from ._runtime import compute_device, stream_frames, upload


class PostChain:
    def __init__(self, grain=None, colormatch=None, lut=None, stencil=None, post_grain=None, device=None):
        """
        grain / post_grain: dict(intensity, saturation_mix, seed, seed_mode=SEED_PER_CLIP)
        colormatch:         dict(reference_image=[1,H,W,3] tensor  |  ref_sums=[1,7] float64, strength)
        lut:                dict(lut_data={"lut","domain_min","domain_max"}, strength 0..10)
        stencil:            dict(op=STENCIL_*, strength, border=BORDER_REPLICATE)
        """
        self.device = torch.device(device) if device is not None else compute_device()
        self.grain, self.colormatch, self.lut, self.stencil, self.post_grain = grain, colormatch, lut, stencil, post_grain
        self._lut_dev = None
        self._ref_sums = None
        self.timing = None          # set to a list to collect (moments_start, moments_end, apply_start, apply_end) CUDA events per call
        # colour-match schedule (vrgdg_chain_cm_apply): one library call; `split` = the three-call path (statistics, parameters,
        # apply as separate entry points; what `timing` needs), `recompute` / `group_frames`: see include/vrgdg_b200.h
        self.split, self.recompute, self.group_frames, self.serial = False, False, 0, False
        self._scratch = None
        if lut is not None:
            self._lut_dev = ops.pack_lut(lut["lut_data"]["lut"], self.device)
        if colormatch is not None:
            self.set_reference(colormatch.get("reference_image"), colormatch.get("ref_sums"))

    # -- colour-match reference statistics (the only cross-rank quantity, see dist.py) --
    def set_reference(self, reference_image=None, ref_sums=None):
        if ref_sums is not None:
            self._ref_sums = ref_sums.to(self.device, torch.float64).reshape(-1, 7).contiguous()
        elif reference_image is not None:
            self._ref_sums = ops.lab_moments(upload(reference_image, self.device))
        else:
            raise ValueError("colour match needs reference_image or ref_sums")

    def _desc(self, frames, first_frame, keep, ext_noise=None, fused_cm=False):
        d = nv.ChainDesc()
        if self.grain is not None:
            s = float(self.grain["saturation_mix"])
            d.grain_enabled = 1
            d.grain_intensity, d.grain_sat, d.grain_one_minus_sat = float(self.grain["intensity"]), s, 1.0 - s
            d.grain_seed = int(self.grain.get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
            d.grain_frame0 = int(first_frame)
            d.grain_seed_mode = int(self.grain.get("seed_mode", nv.SEED_PER_CLIP))
        if self.lut is not None:
            data = self.lut["lut_data"]
            blend = max(0.0, min(10.0, float(self.lut.get("strength", 10.0)))) / 10.0
            if blend > 0.0:
                fdt = torch.float32 if frames.dtype == torch.uint8 else frames.dtype     # uint8 frames become float32 tensors in the reference
                dmin = data["domain_min"].to(dtype=fdt)
                span = torch.clamp(data["domain_max"].to(dtype=fdt) - dmin, min=1e-6)
                d.lut_enabled = 1
                d.lut = self._lut_dev.data.data_ptr()
                d.lut_size = self._lut_dev.size
                d.lut_dmin = (ctypes.c_float * 3)(*dmin.float().tolist())
                d.lut_dspan = (ctypes.c_float * 3)(*span.float().tolist())
                d.lut_blend, d.lut_one_minus_blend = blend, 1.0 - blend
        if self.stencil is not None and not (self.stencil["op"] == nv.STENCIL_NONE):
            d.stencil_op = int(self.stencil["op"])
            d.stencil_strength = float(self.stencil["strength"])
            d.stencil_border = int(self.stencil.get("border", nv.BORDER_REPLICATE))
        if self.post_grain is not None:
            s = float(self.post_grain["saturation_mix"])
            d.post_grain_enabled = 1
            d.post_intensity, d.post_sat, d.post_one_minus_sat = float(self.post_grain["intensity"]), s, 1.0 - s
            d.post_seed = int(self.post_grain.get("seed", 0)) & 0xFFFFFFFFFFFFFFFF
            d.post_frame0 = int(first_frame)
            d.post_seed_mode = int(self.post_grain.get("seed_mode", nv.SEED_PER_FRAME))
        if self.colormatch is not None and fused_cm:
            t = float(self.colormatch.get("strength", 1.0))
            d.colormatch_enabled = 1
            d.cm_t, d.cm_one_minus_t = t, 1.0 - t
        elif self.colormatch is not None:
            # per-frame LAB moments of the colour-match INPUT (= grain output when grain is enabled): a first pass
            # that recomputes the counter-based grain instead of materialising it
            if self.timing is not None:
                self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                self._ev[0].record()
            sums = ops.chain_lab_moments(frames, d, ext_noise=ext_noise)
            params = ops.colormatch_params(sums, self._ref_sums)
            if self.timing is not None:
                self._ev[1].record()
            keep.append(params)
            t = float(self.colormatch.get("strength", 1.0))
            d.colormatch_enabled = 1
            d.cm_params = params.data_ptr()
            d.cm_t, d.cm_one_minus_t = t, 1.0 - t
        return d

    def __call__(self, frames, first_frame=0, ext_noise=None, out=None, fast_math=False):
        """frames: CUDA [B,H,W,3]; first_frame: absolute index of frames[0] in the clip (keys the grain).
        ext_noise (tests): N(0,1) tensor replacing the generator; fast_math then selects the production arithmetic."""
        keep = []
        if self.colormatch is not None and not self.split and self.timing is None:
            d = self._desc(frames, first_frame, keep, ext_noise, fused_cm=True)
            res, self._scratch = ops.chain_cm_apply(frames, d, self._ref_sums, ext_noise=ext_noise, out=out, fast_math=fast_math,
                                                    recompute=self.recompute, group_frames=self.group_frames, scratch=self._scratch, serial=self.serial)
            return res
        d = self._desc(frames, first_frame, keep, ext_noise)
        if self.timing is None:
            return ops.chain_apply(frames, d, ext_noise=ext_noise, keepalive=keep, out=out, fast_math=fast_math)
        ev = getattr(self, "_ev", None) if self.colormatch is not None else None
        if ev is None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            ev[1].record()
        ev[2].record()
        res = ops.chain_apply(frames, d, ext_noise=ext_noise, keepalive=keep, out=out, fast_math=fast_math)
        ev[3].record()
        self.timing.append(tuple(ev))
        self._ev = None
        return res

    def run_host(self, frames_cpu, chunk_frames=8, first_frame=0, out=None):
        """Host frames in, host frames out: chunked upload / compute / download on three streams.  Pass pinned tensors
        (and a reusable pinned `out`) for asynchronous copies."""
        return stream_frames(frames_cpu, lambda f, i: self(f, first_frame + i), chunk_frames, torch.device("cpu"), self.device, out=out)
