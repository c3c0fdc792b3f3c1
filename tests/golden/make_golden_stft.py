#!/usr/bin/env python
"""Golden vectors of the STFT front / back end: scipy.signal.stft / istft called the way the reference's notebooks call
them (window="hann", nperseg=n_fft, noverlap=n_fft-hop).  Build container only:

    python tests/golden/make_golden_stft.py      # -> tests/golden/stft.npz
"""
import os

import numpy as np
import scipy.signal as ss

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
out = {}
cases = [("a", (2,), 1000, 64, 16), ("b", (3, 2), 4097, 256, 128), ("c", (1,), 700, 512, 128), ("d", (2,), 5000, 1024, 256)]
for name, lead, L, n_fft, hop in cases:
    x = rng.standard_normal(lead + (L,))
    f, t, Z = ss.stft(x, window="hann", nperseg=n_fft, noverlap=n_fft - hop)
    t2, y = ss.istft(Z, window="hann", nperseg=n_fft, noverlap=n_fft - hop)
    Zr = Z * (1.0 + 0.1 * rng.standard_normal(Z.shape)) + 0.05j * rng.standard_normal(Z.shape)  # not a valid STFT
    _, yr = ss.istft(Zr, window="hann", nperseg=n_fft, noverlap=n_fft - hop)
    out.update({name + "_x": x, name + "_n_fft": n_fft, name + "_hop": hop, name + "_f": f, name + "_t": t, name + "_Z": Z,
                name + "_y": y, name + "_ty": t2, name + "_Zr": Zr, name + "_yr": yr})
w = np.sqrt(np.hanning(129)[:128])  # an array window
x = rng.standard_normal((2, 2000))
_, _, Z = ss.stft(x, window=w, nperseg=128, noverlap=96)
_, y = ss.istft(Z, window=w, nperseg=128, noverlap=96)
out.update(w_x=x, w_win=w, w_Z=Z, w_y=y)
np.savez_compressed(os.path.join(HERE, "stft.npz"), **out)
print("wrote", os.path.join(HERE, "stft.npz"), len(out), "arrays")
