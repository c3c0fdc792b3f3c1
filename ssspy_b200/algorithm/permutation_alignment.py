"""Correlation-based permutation solver on the device (host mirror of
ssspy/algorithm/permutation_alignment.py:12-121).

``sequence`` has the reference's layout (n_bins, n_sources, n_frames) [a leading batch axis is allowed]; every
positional argument of shape (n_bins, n_sources, ...) is permuted along its source axis like ``sequence``.  The
per-bin correlations and the greedy alignment over the bins run in libssb.so (``ssb_permutation_correlation`` /
``ssb_permutation_align``); the argsort of the correlations is host logic (``numpy.argsort`` on float64, the
reference's own call).  Permutations are indices: the result is an exact rearrangement of the input.
"""
import functools

import numpy as np
import torch

from .. import _device, _lib
from ..special.flooring import EPS, max_flooring
from ..utils.flooring import flooring_to_enum


def _solve(Yb, flooring_fn):
    """Yb: CUDA complex64 (B, N, I, J), permuted in place.  Returns perms (B, I, N) int64 on the device."""
    mode, eps = flooring_to_enum(flooring_fn)
    B, N, I, J = Yb.shape
    corr = _device.empty((B, I), torch.float64)
    _lib.call("ssb_permutation_correlation", Yb.data_ptr(), corr.data_ptr(), B, N, I, J, mode, eps, _device.stream_ptr())
    order = np.argsort(corr.cpu().numpy(), axis=1).astype(np.int32)
    order_d = torch.from_numpy(np.ascontiguousarray(order)).to(Yb.device)
    perms = _device.empty((B, I, N), torch.int32)
    _lib.call("ssb_permutation_align", Yb.data_ptr(), None, order_d.data_ptr(), perms.data_ptr(), B, N, I, J, mode, eps,
              _device.stream_ptr())
    return perms.to(torch.int64)


def correlation_based_permutation_solver(sequence, *args, flooring_fn=functools.partial(max_flooring, eps=EPS),
                                         overwrite=True):
    is_t = _device.is_tensor(sequence)
    if sequence.ndim not in (3, 4):
        raise AssertionError("Dimension of sequence is expected to be 3.")
    for pos_idx, arg in enumerate(args):
        if tuple(arg.shape[:sequence.ndim - 1]) != tuple(sequence.shape[:sequence.ndim - 1]):
            raise ValueError("The shape of {}th argument is invalid.".format(pos_idx + 1))
    batched = sequence.ndim == 4
    Y = _device.to_device(sequence, torch.complex64)
    Yb = (Y if batched else Y.unsqueeze(0)).permute(0, 2, 1, 3).contiguous()  # (B, N, I, J), always a copy
    perms = _solve(Yb, flooring_fn)                                           # (B, I, N)
    idx = perms if batched else perms[0]

    def permute(a):
        """a: (..., n_bins, n_sources, *rest) -> rows gathered by the permutation of each bin."""
        if _device.is_tensor(a):
            ix = idx.to(a.device).reshape(idx.shape + (1,) * (a.dim() - idx.dim())).expand(a.shape)
            res = torch.gather(a, idx.dim() - 1, ix)
            if overwrite:
                a.copy_(res)
                return a
            return res
        ix = idx.cpu().numpy().reshape(tuple(idx.shape) + (1,) * (a.ndim - idx.dim()))
        res = np.take_along_axis(a, ix, axis=idx.dim() - 1)
        if overwrite:
            a[...] = res
            return a
        return res

    out = permute(sequence)
    others = tuple(permute(a) for a in args)
    if not is_t:
        out = np.asarray(out)
    if len(others) == 0:
        return out
    if len(others) == 1:
        return out, others[0]
    return out, others
