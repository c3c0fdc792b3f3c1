"""Minimal distortion principle on the device (host mirror of
ssspy/algorithm/minimal_distortion_principle.py:6-43)."""
import numpy as np
import torch

from .. import _device, _lib
from ..utils.select_pair import wrap_reference_id


def minimal_distortion_principle(estimated, reference=None, reference_id=0):
    """``estimated`` (N, I, J) [or (B, N, I, J)], ``reference`` of the same shape ->
    ``conj(z) * estimated`` with ``z[n,i] = sum_j y conj(x_ref) / sum_j |y|^2``.  ``reference_id=None``
    returns every reference channel on a new leading axis, as the reference does."""
    is_t = _device.is_tensor(estimated)
    Y = _device.to_device(estimated, torch.complex64)
    X = _device.to_device(reference, torch.complex64)
    batched = Y.dim() == 4
    Yb = Y if batched else Y.unsqueeze(0)
    Xb = X if batched else X.unsqueeze(0)
    B, N, I, J = Yb.shape
    outs = []
    for ref in (range(Xb.shape[1]) if reference_id is None else [reference_id]):
        out = torch.empty_like(Yb)
        _lib.call("ssb_minimal_distortion_principle", Yb.data_ptr(), Xb.data_ptr(), out.data_ptr(), B, N, I, J,
                  wrap_reference_id(ref, Xb.shape[1]),
                  _device.stream_ptr())
        outs.append(out if batched else out[0])
    res = torch.stack(outs, dim=0) if reference_id is None else outs[0]
    return res if is_t else res.cpu().numpy().astype(np.complex128)
