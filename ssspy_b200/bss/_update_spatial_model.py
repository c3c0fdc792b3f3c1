"""Standalone spatial-update operators on the device (host mirror of
ssspy/bss/_update_spatial_model.py: update_by_ip1 :17-78, update_by_ip2 :81-143,
update_by_ip2_one_pair :317-395, update_by_iss1 :146-194, update_by_iss2 :197-314, update_by_ipa :398-513).

Same argument meaning as the reference; leading batch axes are allowed.  NumPy in -> NumPy
(complex128) out, CUDA tensors in -> CUDA tensors out.  ``overwrite=True`` writes the result back
into the array that was passed, as the reference does.
"""
import functools

import numpy as np
import torch

from .. import _device, _lib
from ..special.flooring import EPS, max_flooring
from ..utils.flooring import flooring_to_enum
from ..utils.select_pair import sequential_pair_selector, wrap_pairs

_DEFAULT_FLOOR = functools.partial(max_flooring, eps=EPS)


def _finish(res, orig, overwrite):
    _lib.check_status()  # LinAlgError("Singular matrix") where the reference's np.linalg.solve raises (_solve.py:15)
    if _device.is_tensor(orig):
        if overwrite:
            orig.copy_(res.to(orig.dtype))
            return orig
        return res.to(orig.dtype)
    out = res.cpu().numpy().astype(np.result_type(orig.dtype, np.complex64) if np.iscomplexobj(orig) else np.complex128)
    if overwrite:
        orig[...] = out
        return orig
    return out


def update_by_ip1(demix_filter, weighted_covariance, flooring_fn=_DEFAULT_FLOOR, overwrite=True):
    """W (*, n_bins, N, N), U (*, n_bins, N, N, N) -> updated W (Gauss-Seidel over sources)."""
    mode, eps = flooring_to_enum(flooring_fn)
    W = _device.to_device(demix_filter, torch.complex64).clone()
    U = _device.to_device(weighted_covariance, torch.complex64)
    N = W.shape[-1]
    assert tuple(U.shape[-3:]) == (N, N, N) and U.shape[:-3] == W.shape[:-2], "shape mismatch of W and U"
    _lib.call("ssb_update_by_ip1", W.data_ptr(), U.data_ptr(), W.numel() // (N * N), N, mode, eps,
              _device.stream_ptr())
    return _finish(W, demix_filter, overwrite)


def update_by_ip2(demix_filter, weighted_covariance, flooring_fn=_DEFAULT_FLOOR, pair_selector=None, overwrite=True):
    mode, eps = flooring_to_enum(flooring_fn)
    if pair_selector is None:
        pair_selector = sequential_pair_selector
    W = _device.to_device(demix_filter, torch.complex64).clone()
    U = _device.to_device(weighted_covariance, torch.complex64)
    N = W.shape[-1]
    pairs = wrap_pairs(pair_selector(N), N)
    for s in range(0, len(pairs), _lib.SSB_MAX_PAIRS):
        chunk = pairs[s:s + _lib.SSB_MAX_PAIRS]
        _lib.call("ssb_update_by_ip2", W.data_ptr(), U.data_ptr(), W.numel() // (N * N), N, _lib.pairs_array(chunk),
                  len(chunk), mode, eps, _device.stream_ptr())
    return _finish(W, demix_filter, overwrite)


def update_by_ip2_one_pair(demix_filter, weighted_covariance_pair, pair, flooring_fn=_DEFAULT_FLOOR):
    """Returns the updated pair of rows, shape (*, n_bins, 2, N) (the reference does not write W)."""
    mode, eps = flooring_to_enum(flooring_fn)
    W = _device.to_device(demix_filter, torch.complex64).clone()
    U = _device.to_device(weighted_covariance_pair, torch.complex64)
    N = W.shape[-1]
    (m, n), = wrap_pairs([pair], N)
    _lib.call("ssb_update_by_ip2_one_pair", W.data_ptr(), U.data_ptr(), W.numel() // (N * N), N, m, n, mode, eps,
              _device.stream_ptr())
    res = W[..., (m, n), :]
    if _device.is_tensor(demix_filter):
        return res.to(demix_filter.dtype)
    return res.cpu().numpy().astype(np.complex128)


def update_by_iss1(separated, weight, flooring_fn=_DEFAULT_FLOOR):
    """Y (N, I, J) [or (B, N, I, J)], weight broadcastable to Y's shape (e.g. (N, 1, J) for IVA)."""
    mode, eps = flooring_to_enum(flooring_fn)
    Y = _device.to_device(separated, torch.complex64).clone()
    batched = Y.dim() == 4
    Yb = Y if batched else Y.unsqueeze(0)
    B, N, I, J = Yb.shape
    phi = _device.to_device(weight, torch.float32)
    phi = phi if batched else phi.unsqueeze(0)
    if phi.shape[-2] == 1:  # per-frame weights shared by all bins
        phi = phi.expand(B, N, 1, J).contiguous()
        sb, sn, si = N * J, J, 0
    else:
        phi = phi.expand(B, N, I, J).contiguous()
        sb, sn, si = N * I * J, I * J, J
    _lib.call("ssb_update_by_iss1", Yb.data_ptr(), phi.data_ptr(), sb, sn, si, B, N, I, J, mode, eps,
              _device.stream_ptr())
    if _device.is_tensor(separated):
        return Y.to(separated.dtype)
    return Y.cpu().numpy().astype(np.complex128)


def update_by_iss2(separated, weight, flooring_fn=_DEFAULT_FLOOR, pair_selector=None):
    """Pairwise iterative source steering (ssspy/bss/_update_spatial_model.py:197-314).  ``separated`` (N, I, J) [or
    (B, N, I, J)], ``weight`` broadcastable to its shape; the default selector yields (0, 1), (2, 3), ...
    (``sequential_pair_selector(N, stop=N, step=2)``, :233-234); negative indices wrap as in the reference."""
    mode, eps = flooring_to_enum(flooring_fn)
    Y = _device.to_device(separated, torch.complex64).clone()
    batched = Y.dim() == 4
    Yb = Y if batched else Y.unsqueeze(0)
    B, N, I, J = Yb.shape
    if pair_selector is None:
        pair_selector = functools.partial(sequential_pair_selector, stop=N, step=2)
    pairs = wrap_pairs(pair_selector(N), N)
    phi = _device.to_device(weight, torch.float32)
    phi = phi if batched else phi.unsqueeze(0)
    if phi.shape[-2] == 1:
        phi = phi.expand(B, N, 1, J).contiguous()
        sb, sn, si = N * J, J, 0
    else:
        phi = phi.expand(B, N, I, J).contiguous()
        sb, sn, si = N * I * J, I * J, J
    for q0 in range(0, len(pairs), _lib.SSB_MAX_PAIRS):
        chunk = pairs[q0:q0 + _lib.SSB_MAX_PAIRS]
        _lib.call("ssb_update_by_iss2", Yb.data_ptr(), phi.data_ptr(), sb, sn, si, B, N, I, J,
                  _lib.pairs_array(chunk), len(chunk), mode, eps, _device.stream_ptr())
    if _device.is_tensor(separated):
        return Y.to(separated.dtype)
    return Y.cpu().numpy().astype(np.complex128)


def update_by_ipa(separated, weight, normalization=True, flooring_fn=_DEFAULT_FLOOR, max_iter=1):
    """Iterative projection with adjustment (ssspy/bss/_update_spatial_model.py:398-513): ``separated`` (N, I, J) [or
    (B, N, I, J)], ``weight`` broadcastable to its shape, ``normalization`` / ``max_iter`` as in the reference (trace
    normalisation of the LQPQM and number of Newton-Raphson updates, ssspy/linalg/lqpqm.py:13-219)."""
    mode, eps = flooring_to_enum(flooring_fn)
    Y = _device.to_device(separated, torch.complex64).clone()
    batched = Y.dim() == 4
    Yb = Y if batched else Y.unsqueeze(0)
    B, N, I, J = Yb.shape
    phi = _device.to_device(weight, torch.float32)
    phi = phi if batched else phi.unsqueeze(0)
    if phi.shape[-2] == 1 and I != 1:
        phi = phi.expand(B, N, 1, J).contiguous()
        sb, sn, si = N * J, J, 0
    else:
        phi = phi.expand(B, N, I, J).contiguous()
        sb, sn, si = N * I * J, I * J, J
    _lib.call("ssb_update_by_ipa", Yb.data_ptr(), phi.data_ptr(), sb, sn, si, B, N, I, J, int(bool(normalization)),
              int(max_iter), mode, eps, _device.stream_ptr())
    if _device.is_tensor(separated):
        return Y.to(separated.dtype)
    return Y.cpu().numpy().astype(np.complex128)
