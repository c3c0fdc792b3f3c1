"""CPU restatement of the STFT conventions of the reference's notebooks (scipy.signal.stft / istft with window="hann",
nperseg=n_fft, noverlap=n_fft-hop; notebooks/BSS/ILRMA/GaussILRMA-IP1-MM.ipynb) in plain NumPy.  Test infrastructure
only (see oracle/__init__.py); pinned against scipy's own output in tests/golden/stft.npz."""
import numpy as np


def hann(nperseg):
    n = np.arange(nperseg)
    return 0.5 - 0.5 * np.cos(2 * np.pi * n / nperseg)


def stft(x, window, hop):
    """x (*, L) real -> (*, nperseg // 2 + 1, n_frames): zero extension by nperseg // 2, zero padding to whole hops,
    rfft of the windowed segments / sum(window)  (scipy.signal._spectral_py._spectral_helper, mode='stft')."""
    x = np.asarray(x, dtype=np.float64)
    n = len(window)
    ext = np.concatenate([np.zeros(x.shape[:-1] + (n // 2,)), x, np.zeros(x.shape[:-1] + (n // 2,))], axis=-1)
    nadd = (-(ext.shape[-1] - n) % hop) % n
    ext = np.concatenate([ext, np.zeros(x.shape[:-1] + (nadd,))], axis=-1)
    n_frames = (ext.shape[-1] - (n - hop)) // hop
    seg = np.stack([ext[..., f * hop:f * hop + n] for f in range(n_frames)], axis=-2)  # (*, frames, n)
    Z = np.fft.rfft(seg * window, axis=-1) / window.sum()
    return np.swapaxes(Z, -1, -2)


def istft(Z, window, hop):
    """(*, nperseg // 2 + 1, n_frames) -> (*, nperseg + (n_frames - 1) hop - 2 (nperseg // 2)): weighted overlap-add
    (scipy.signal.istft)."""
    n = len(window)
    n_frames = Z.shape[-1]
    seg = np.fft.irfft(np.swapaxes(Z, -1, -2), n=n, axis=-1) * window.sum() * window
    out = np.zeros(Z.shape[:-2] + (n + (n_frames - 1) * hop,))
    norm = np.zeros(n + (n_frames - 1) * hop)
    for f in range(n_frames):
        out[..., f * hop:f * hop + n] += seg[..., f, :]
        norm[f * hop:f * hop + n] += window ** 2
    out, norm = out[..., n // 2:out.shape[-1] - n // 2], norm[n // 2:len(norm) - n // 2]
    return out / np.where(norm > 1e-10, norm, 1.0)
