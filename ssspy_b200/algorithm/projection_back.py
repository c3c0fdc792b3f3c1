"""Projection back on the device (host mirror of ssspy/algorithm/projection_back.py:6-121)."""
import numpy as np
import torch

from .. import _device, _lib
from ..utils.select_pair import wrap_reference_id


def projection_back(data_or_filter, reference=None, reference_id=0):
    """Filter form (``reference is None``): ``W (*, N, N) -> W * (W^-1)[..., ref, :, None]``
    (projection_back.py:87-99).  Spectrogram form: ``Y (N, I, J)`` [or ``(B, N, I, J)``], reference
    ``X`` of the same shape -> ``Y`` scaled by ``(X Y^H (Y Y^H)^-1)[ref, n]`` (:100-121).
    ``reference_id=None`` returns every reference channel on a new leading axis."""
    is_t = _device.is_tensor(data_or_filter)
    st = _device.stream_ptr
    if reference is None:
        W = _device.to_device(data_or_filter, torch.complex64)
        N = W.shape[-1]
        n_mat = W.numel() // (N * N)
        refs = range(N) if reference_id is None else [reference_id]
        outs = []
        for ref in refs:
            out = torch.empty_like(W)
            _lib.call("ssb_projection_back_w", W.data_ptr(), out.data_ptr(), n_mat, N, wrap_reference_id(ref, N), st())
            outs.append(out)
        res = torch.stack(outs, dim=0) if reference_id is None else outs[0]
    else:
        Y = _device.to_device(data_or_filter, torch.complex64)
        X = _device.to_device(reference, torch.complex64)
        batched = Y.dim() == 4
        Yb = Y if batched else Y.unsqueeze(0)
        Xb = X if batched else X.unsqueeze(0)
        B, N, I, J = Yb.shape
        scale = torch.empty((B, I, N, N), dtype=torch.complex64, device=Yb.device)
        refs = range(N) if reference_id is None else [reference_id]
        outs = []
        for ref in refs:
            out = torch.empty_like(Yb)
            _lib.call("ssb_projection_back_y", Yb.data_ptr(), Xb.data_ptr(), out.data_ptr(), scale.data_ptr(),
                      B, N, I, J, wrap_reference_id(ref, N), st())
            outs.append(out if batched else out[0])
        res = torch.stack(outs, dim=0) if reference_id is None else outs[0]
    _lib.check_status()  # LinAlgError("Singular matrix") like np.linalg.inv in the reference (projection_back.py:89, :110)
    if is_t:
        return res
    return res.cpu().numpy().astype(np.complex128)
