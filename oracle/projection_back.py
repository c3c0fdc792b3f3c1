"""Projection back, W-form and Y-form (oracle; see oracle/__init__.py).

ssspy/algorithm/projection_back.py:87-99 (demixing-filter form) and :100-121
(spectrogram form)."""
import numpy as np


def projection_back(data_or_filter, reference=None, reference_id=0):
    if reference is None:
        W = data_or_filter  # (*, N, N)
        scale = np.linalg.inv(W)
        if reference_id is None:
            scale = np.moveaxis(scale[..., np.newaxis], -3, 0)  # (N_ch, *, N_src, 1)
            return W * scale
        return W * scale[..., reference_id, :][..., np.newaxis]
    Y = data_or_filter.transpose(1, 0, 2)  # (I, N, J)
    X = reference.transpose(1, 0, 2)
    YH = np.conj(Y.transpose(0, 2, 1))
    scale = (X @ YH) @ np.linalg.inv(Y @ YH)  # (I, N_ch, N_src)
    if reference_id is None:
        scale = scale.transpose(1, 0, 2)  # (N_ch, I, N_src)
        return (Y * scale[..., np.newaxis]).swapaxes(-3, -2)  # (N_ch, N_src, I, J)
    scale = scale[..., reference_id, :]
    return (Y * scale[..., np.newaxis]).swapaxes(-3, -2)
