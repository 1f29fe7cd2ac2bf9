"""Device policy and host<->device frame streaming shared by the node classes."""
import torch


def _comfy_mm():
    try:
        import comfy.model_management as mm  # provided by the ComfyUI host process
        return mm
    except Exception:
        return None


def compute_device(hint=None):
    """The CUDA device the kernels run on.  Mirrors comfy.model_management.get_torch_device() (nodes.py:42);
    outside ComfyUI: the tensor's own CUDA device, else the current CUDA device.  Never a CPU."""
    if isinstance(hint, torch.Tensor) and hint.device.type == "cuda":
        return hint.device
    mm = _comfy_mm()
    if mm is not None:
        dev = torch.device(mm.get_torch_device())
        if dev.type == "cuda":
            return dev
    if not torch.cuda.is_available():
        raise RuntimeError("vrgdg_b200: no CUDA device is available; these nodes are sm_100a kernels and have no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def result_device(images, numpy_path=False):
    """Where a node returns its IMAGE.  Inside ComfyUI: exactly what the reference does
    (intermediate_device() nodes.py:65,123,177; CPU for the numpy paths :209).  Outside: the input's device."""
    mm = _comfy_mm()
    if mm is not None:
        return torch.device("cpu") if numpy_path else torch.device(mm.intermediate_device())
    return images.device


def stream_frames(src, fn, chunk, out_device, device=None, out=None):
    """Apply fn(cuda_frames, first_frame_index) -> cuda_frames over src [B,...] in chunks of `chunk` frames.

    CUDA input: a single call (chunk ignored).  CPU input: chunks are uploaded on a side stream while the previous
    chunk computes, results are downloaded on a third stream; pinned source/result tensors make those copies truly
    asynchronous (result is pinned when the source is).  `out`: optional preallocated host result (reused across calls so
    that pinning cost is paid once)."""
    B = int(src.shape[0])
    out_device = torch.device(out_device)
    if src.device.type == "cuda":
        res = fn(src, 0)
        return res if res.device == out_device else res.to(out_device)
    dev = device if device is not None else compute_device()
    if B == 0:
        return torch.empty_like(src, device=out_device)
    chunk = B if chunk is None or chunk <= 0 else int(chunk)
    src = src.contiguous()
    to_cpu = out_device.type == "cpu"
    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        if out is None:
            out = torch.empty(src.shape, dtype=src.dtype, pin_memory=src.is_pinned()) if to_cpu else torch.empty(src.shape, dtype=src.dtype, device=out_device)
        elif out.shape != src.shape or out.dtype != src.dtype or out.device != out_device:
            raise ValueError("vrgdg_b200: `out` must match the source frames in shape and dtype and live on %s" % out_device)
        pending = None
        for i in range(0, B, chunk):
            with torch.cuda.stream(up):
                d_in = src[i:i + chunk].to(dev, non_blocking=True)
                ev_up = torch.cuda.Event()
                ev_up.record(up)
            compute.wait_event(ev_up)
            d_in.record_stream(compute)
            d_out = fn(d_in, i)
            ev_done = torch.cuda.Event()
            ev_done.record(compute)
            down.wait_event(ev_done)
            with torch.cuda.stream(down):
                d_out.record_stream(down)
                out[i:i + chunk].copy_(d_out, non_blocking=True)
            pending = d_out
        down.synchronize()
        del pending
    return out
