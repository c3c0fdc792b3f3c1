"""Device-buffer plumbing: PyTorch is used only as the allocator / stream provider for the CUDA
library (no torch op is on the compute path)."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available() or _lib.device_count() == 0:
        raise RuntimeError(
            "ssspy_b200 needs a CUDA device (sm_100a): there is no CPU fallback for the demixing path.")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def to_device(x, dtype):
    """NumPy array or tensor -> contiguous CUDA tensor of ``dtype`` (copy unless already one)."""
    dev = require_cuda()
    if is_tensor(x):
        if x.is_cuda and x.dtype == dtype and x.is_contiguous():
            return x  # zero-copy
        return x.to(device=dev, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(x)
    if np.iscomplexobj(arr):
        arr = arr.astype(np.complex64 if dtype == torch.complex64 else np.complex128, copy=False)
    else:
        arr = arr.astype({torch.float32: np.float32, torch.float64: np.float64,
                          torch.complex64: np.complex64, torch.complex128: np.complex128}[dtype], copy=False)
    return torch.from_numpy(arr).to(dev)


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def ptr(t):
    return 0 if t is None else t.data_ptr()


def to_host(t, dtype=None):
    a = t.detach().cpu().numpy()
    return a if dtype is None else a.astype(dtype)
