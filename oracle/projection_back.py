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


def minimal_distortion_principle(estimated, reference, reference_id=0):
    """z = sum_j y conj(x_ref) / sum_j |y|^2; output = conj(z) y
    (ssspy/algorithm/minimal_distortion_principle.py:31-43)."""
    Y, Xc = estimated, np.conj(reference)
    if reference_id is None:
        num = np.sum(Y * Xc[:, np.newaxis, :, :], axis=-1, keepdims=True)
    else:
        num = np.sum(Y * Xc[reference_id], axis=-1, keepdims=True)
    return np.conj(num / np.sum(np.abs(Y) ** 2, axis=-1, keepdims=True)) * Y
