"""Device-memory plumbing: torch owns allocations and streams, the kernels are ours.

torch is used only to allocate device buffers, copy host<->device and expose the
current CUDA stream; every kernel launched from this package lives in
libradnet_b200.so.
"""
import ctypes

import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "rock_art_radnet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def is_cuda_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Raw device (or host, for numpy) pointer as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


_TORCH_OF = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
    np.dtype(np.uint8): torch.uint8,
}


def to_device(x, dtype, device):
    """numpy / torch input -> contiguous CUDA tensor of `dtype` (numpy dtype) on `device`."""
    tdt = _TORCH_OF[np.dtype(dtype)]
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=tdt).contiguous()
    arr = np.ascontiguousarray(np.asarray(x), dtype=dtype)
    return torch.from_numpy(arr).to(device)


def empty(shape, dtype, device):
    return torch.empty(shape, dtype=_TORCH_OF[np.dtype(dtype)], device=device)


def zeros(shape, dtype, device):
    return torch.zeros(shape, dtype=_TORCH_OF[np.dtype(dtype)], device=device)


def host_f64(values):
    """Small host-side float64 table handed to an `h_` parameter."""
    return np.ascontiguousarray(np.asarray(values, dtype=np.float64))
