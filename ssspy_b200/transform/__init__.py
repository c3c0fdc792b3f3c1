"""STFT / inverse STFT on the device: the step either side of the separators.

ssspy has no transform of its own: its notebooks call ``scipy.signal.stft(waveform, window="hann", nperseg=n_fft,
noverlap=n_fft - hop_length)`` and ``scipy.signal.istft`` with the same arguments
(notebooks/BSS/ILRMA/GaussILRMA-IP1-MM.ipynb).  ``stft`` / ``istft`` below take those arguments, follow scipy's defaults
(``boundary="zeros"``, ``padded=True``, one-sided spectrum, ``scaling="spectrum"``) and return the same tuples, so a
notebook switches by changing the import.  NumPy in -> NumPy out; CUDA tensors in -> CUDA tensors out (the spectrogram
can then go straight into a separator without a host round trip).  fp64 arithmetic, radix-2 FFT in shared memory
(``ssb_stft.cu``): ``nperseg`` must be a power of two in [16, 8192]; anything scipy offers beyond the options above
raises ``NotImplementedError`` (there is no CPU fallback).
"""
import ctypes

import numpy as np
import torch

from .. import _device, _lib

__all__ = ["stft", "istft"]


def _window(window, nperseg):
    if isinstance(window, str):
        if window not in ("hann", "hanning"):
            raise NotImplementedError("window={!r}: pass 'hann' or an array of nperseg samples".format(window))
        n = np.arange(nperseg, dtype=np.float64)
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nperseg)  # scipy.signal.get_window("hann", nperseg): periodic
    else:
        win = np.asarray(window.cpu() if _device.is_tensor(window) else window, dtype=np.float64)
        if win.ndim != 1:
            raise ValueError("window must be 1-D")
        if win.shape[0] != nperseg:
            raise ValueError("window must have length of nperseg")
    return win


def _hop(nperseg, noverlap):
    nperseg = int(nperseg)
    if nperseg < 1:
        raise ValueError("nperseg must be a positive integer")
    noverlap = nperseg // 2 if noverlap is None else int(noverlap)
    if noverlap >= nperseg:
        raise ValueError("noverlap must be less than nperseg.")
    return nperseg, noverlap, nperseg - noverlap


def _unsupported(**kw):
    for name, (value, default) in kw.items():
        if value != default:
            raise NotImplementedError("{}={!r} is not supported on the device (only {!r})".format(name, value, default))


def stft(x, fs=1.0, window="hann", nperseg=256, noverlap=None, nfft=None, detrend=False, return_onesided=True,
         boundary="zeros", padded=True, axis=-1, scaling="spectrum"):
    """``scipy.signal.stft`` on the device -> ``(f, t, Zxx)`` with ``Zxx`` of shape ``(*, nperseg // 2 + 1, n_frames)``
    complex128."""
    nperseg, noverlap, hop = _hop(nperseg, noverlap)
    _unsupported(nfft=(nperseg if nfft is None else int(nfft), nperseg), detrend=(detrend, False),
                 return_onesided=(return_onesided, True), boundary=(boundary, "zeros"), padded=(padded, True),
                 axis=(axis, -1), scaling=(scaling, "spectrum"))
    is_t = _device.is_tensor(x)
    if (x.is_complex() if is_t else np.iscomplexobj(x)):
        raise NotImplementedError("complex input is not supported (one-sided transform of real signals only)")
    win = _window(window, nperseg)
    xd = _device.to_device(x, torch.float64).contiguous()
    lead, n_samples = tuple(xd.shape[:-1]), int(xd.shape[-1])
    if n_samples < nperseg:  # scipy shrinks nperseg with a warning; a silent change of the bin count helps nobody here
        raise ValueError("nperseg = {} is greater than input length = {}".format(nperseg, n_samples))
    n_rows = int(np.prod(lead)) if lead else 1
    n_frames = ctypes.c_int(0)
    _lib.call("ssb_stft_frames", n_samples, nperseg, hop, ctypes.byref(n_frames))
    n_frames, n_bins = n_frames.value, nperseg // 2 + 1
    wd = _device.to_device(win, torch.float64)
    Z = torch.empty(lead + (n_bins, n_frames), dtype=torch.complex128, device=xd.device)
    _lib.call("ssb_stft", xd.data_ptr(), wd.data_ptr(), float(win.sum()), Z.data_ptr(), n_rows, n_samples, nperseg, hop,
              _device.stream_ptr())
    f = np.fft.rfftfreq(nperseg, 1.0 / fs)
    t = np.arange(nperseg / 2, nperseg / 2 + n_frames * hop, hop)[:n_frames] / float(fs) - (nperseg // 2) / float(fs)
    return f, t, (Z if is_t else Z.cpu().numpy())


def istft(Zxx, fs=1.0, window="hann", nperseg=None, noverlap=None, nfft=None, input_onesided=True, boundary=True,
          time_axis=-1, freq_axis=-2, scaling="spectrum"):
    """``scipy.signal.istft`` on the device -> ``(t, x)`` with ``x`` of shape ``(*, n_samples)`` float64, where
    ``n_samples = nperseg + (n_frames - 1) * hop - 2 * (nperseg // 2)`` (the zero-padded tail of ``stft`` included, as
    in scipy)."""
    is_t = _device.is_tensor(Zxx)
    if Zxx.ndim < 2:
        raise ValueError("Input stft must be at least 2d!")
    n_bins, n_frames = int(Zxx.shape[-2]), int(Zxx.shape[-1])
    if nperseg is None:
        nperseg = 2 * (n_bins - 1)
    nperseg, noverlap, hop = _hop(nperseg, noverlap)
    _unsupported(nfft=(nperseg if nfft is None else int(nfft), nperseg), input_onesided=(input_onesided, True),
                 boundary=(bool(boundary), True), time_axis=(time_axis, -1), freq_axis=(freq_axis, -2),
                 scaling=(scaling, "spectrum"))
    if n_bins != nperseg // 2 + 1:
        raise ValueError("nperseg={} does not match the {} frequency bins of the input".format(nperseg, n_bins))
    win = _window(window, nperseg)
    Zd = _device.to_device(Zxx, torch.complex128).contiguous()
    lead = tuple(Zd.shape[:-2])
    n_rows = int(np.prod(lead)) if lead else 1
    n_out = nperseg + (n_frames - 1) * hop - 2 * (nperseg // 2)
    wd = _device.to_device(win, torch.float64)
    y = torch.empty(lead + (max(n_out, 0),), dtype=torch.float64, device=Zd.device)
    seg = torch.empty((n_rows, n_frames, nperseg), dtype=torch.float64, device=Zd.device)
    _lib.call("ssb_istft", Zd.data_ptr(), wd.data_ptr(), float(win.sum()), y.data_ptr(), seg.data_ptr(), n_rows,
              n_frames, nperseg, hop, _device.stream_ptr())
    t = np.arange(max(n_out, 0)) / float(fs)
    return t, (y if is_t else y.cpu().numpy())
