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
    want = {torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64,
            torch.complex128: np.complex128}[dtype]
    if np.iscomplexobj(arr) and not np.issubdtype(want, np.complexfloating):
        raise TypeError("complex array given where a real one is expected")
    if arr.dtype != want and arr.nbytes >= (1 << 20) and arr.dtype in (np.float64, np.complex128, np.float32):
        # large precision change (float64 state injected by the caller): upload as it is and convert on the device;
        # numpy's single-threaded astype costs more than the extra PCIe bytes (15 ms vs 3 ms for config 2's T, V)
        return torch.from_numpy(arr).to(dev).to(dtype)
    return torch.from_numpy(arr.astype(want, copy=False)).to(dev)


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def ptr(t):
    return 0 if t is None else t.data_ptr()


def to_host(t, dtype=None):
    a = t.detach().cpu().numpy()
    return a if dtype is None else a.astype(dtype)
