"""Device-buffer plumbing: torch tensors as HBM buffers, nothing else.

torch is used for allocation, host<->device copies (pinned staging) and stream
handles; every computation happens in libxrs.so.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import XrsError

_TORCH_DTYPES = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8,
    np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
    np.dtype(np.int64): torch.int64, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32,
}


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise XrsError("no CUDA device available: xcube_resampling_b200 has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def torch_dtype(np_dtype) -> torch.dtype:
    dt = np.dtype(np_dtype)
    if dt not in _TORCH_DTYPES:
        raise TypeError(f"unsupported data type {dt}")
    return _TORCH_DTYPES[dt]


def to_device(array, device=None, dtype=None) -> torch.Tensor:
    """numpy (or torch) array -> contiguous device tensor (H2D through pinned memory)."""
    dev = require_cuda(device)
    if isinstance(array, torch.Tensor):
        t = array
        if dtype is not None:
            t = t.to(torch_dtype(dtype))
        return t.to(dev, non_blocking=True).contiguous()
    a = np.ascontiguousarray(array if dtype is None else np.asarray(array, dtype=dtype))
    if not a.flags.writeable:
        a = a.copy()
    # Pinned inputs (see pinned_empty) go over DMA asynchronously at link speed; pageable ones are
    # staged by the driver.  copy_ into a preallocated device tensor is used instead of Tensor.to():
    # the latter was measured to block and to run at pageable speed (17 GB/s) even for pinned memory
    # wrapped by from_numpy.
    out = torch.empty(a.shape, dtype=torch_dtype(a.dtype), device=dev)
    out.copy_(torch.from_numpy(a), non_blocking=True)
    return out


def to_device_pitched(array, device=None, multiple_bytes: int = 128) -> torch.Tensor:
    """Upload an (.., h, w) array into a device buffer whose row pitch is a multiple of
    ``multiple_bytes`` and return the (.., h, w) view.  A 16-byte aligned pitch is what lets the
    gather kernel stage source boxes with TMA tensor copies (include/xrs.h, xrs_gather_ij)."""
    dev = require_cuda(device)
    if isinstance(array, torch.Tensor):
        src = array
    else:
        a = np.ascontiguousarray(array)
        if not a.flags.writeable:
            a = a.copy()
        src = torch.from_numpy(a)
    w = src.shape[-1]
    per = max(1, multiple_bytes // src.element_size())
    wp = -(-w // per) * per
    buf = torch.empty(tuple(src.shape[:-1]) + (wp,), dtype=src.dtype, device=dev)
    view = buf[..., :w]
    view.copy_(src, non_blocking=True)
    return view


def to_host(t: torch.Tensor) -> np.ndarray:
    """device tensor -> numpy via a pinned host buffer."""
    if t.device.type != "cuda":
        return t.numpy()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=t.numel() * t.element_size() >= (1 << 20))
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def pinned_empty(shape, np_dtype) -> np.ndarray:
    """Page-locked host array (numpy view of a pinned torch tensor) for fast H2D/D2H."""
    t = torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(np_dtype), pin_memory=torch.cuda.is_available())
    return t.numpy()


def empty(shape, np_dtype, device=None) -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(np_dtype), device=require_cuda(device))


def workspace(n_bytes: int, device=None) -> torch.Tensor:
    return torch.empty(max(int(n_bytes), 16), dtype=torch.uint8, device=require_cuda(device))


def ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr_array(tensors) -> ctypes.Array:
    arr = (ctypes.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def plane_ptr_array(t: torch.Tensor) -> ctypes.Array:
    """Pointers to the planes t[0], t[1], ... of a (bands, h, w) tensor, computed from the base
    pointer and the band stride (no per-band tensor views: this sits on the per-call hot path)."""
    n = t.shape[0]
    base, step = t.data_ptr(), t.stride(0) * t.element_size()
    arr = (ctypes.c_void_p * n)()
    for k in range(n):
        arr[k] = base + k * step
    return arr


class BandPipeline:
    """Host -> device -> host streaming of a (bands, h, w) variable in band chunks.

    Three streams: uploads (H2D) of chunk k+1, the kernels of chunk k on the caller's current stream
    and the download (D2H) of chunk k-1 run concurrently, so a PCIe-bound call costs
    max(H2D, D2H) instead of their sum.  Two device slots per direction (double buffering); the
    source slots are row-pitched to 128 bytes (TMA-addressable, see :func:`to_device_pitched`).
    ``process(src_chunk, out_chunk)`` must only enqueue work on the current stream.
    """

    def __init__(self, values: np.ndarray, out_shape_hw, out_dtype, device=None, chunk_bands: int = 4,
                 pitch_bytes: int = 128):
        self.dev = require_cuda(device)
        v = np.ascontiguousarray(values)
        if not v.flags.writeable:
            v = v.copy()
        self.src_host = torch.from_numpy(v)
        self.bands, self.h, self.w = v.shape
        self.chunk = max(1, min(int(chunk_bands), self.bands))
        per = max(1, pitch_bytes // self.src_host.element_size())
        wp = -(-self.w // per) * per
        self.in_slots = [torch.empty((self.chunk, self.h, wp), dtype=self.src_host.dtype, device=self.dev)
                         for _ in range(2)]
        H, W = int(out_shape_hw[0]), int(out_shape_hw[1])
        self.out_slots = [torch.empty((self.chunk, H, W), dtype=torch_dtype(out_dtype), device=self.dev)
                          for _ in range(2)]
        self.out_host = pinned_empty((self.bands, H, W), out_dtype)
        self._out_host_t = torch.from_numpy(self.out_host)

    def run(self, process) -> np.ndarray:
        main = torch.cuda.current_stream(self.dev)
        s_in, s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]
        out_ready = [torch.cuda.Event() for _ in range(2)]
        out_free = [torch.cuda.Event() for _ in range(2)]
        for e in in_free + out_free:
            e.record(main)
        # chunk schedule 1, 2, 4, 4, ...: a short first chunk lets the download engine start early
        starts, b0, size = [], 0, 1
        while b0 < self.bands:
            nb = min(size, self.chunk, self.bands - b0)
            starts.append((b0, nb))
            b0 += nb
            size *= 2
        for k, (b0, nb) in enumerate(starts):
            slot = k % 2
            src_view = self.in_slots[slot][:nb, :, : self.w]
            out_view = self.out_slots[slot][:nb]
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[slot])
                src_view.copy_(self.src_host[b0:b0 + nb], non_blocking=True)
                in_ready[slot].record(s_in)
            main.wait_event(in_ready[slot])
            main.wait_event(out_free[slot])
            process(src_view, out_view)
            in_free[slot].record(main)
            out_ready[slot].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(out_ready[slot])
                self._out_host_t[b0:b0 + nb].copy_(out_view, non_blocking=True)
                out_free[slot].record(s_out)
        s_out.synchronize()
        main.synchronize()
        return self.out_host
