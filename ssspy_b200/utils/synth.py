"""Synthetic STFT mixtures for parity tests and the benchmark.

No audio is needed: per mixture ``b`` of config ``c`` the generator is
``np.random.default_rng(1000*c + b)``.  Two modes:

* ``iid``  -- i.i.d. complex Gaussian ``X`` (what the reference's docstrings use,
  ssspy/bss/ilrma.py:655-657);
* ``mix``  -- low-rank-variance sources through a random per-bin mixing matrix
  (well conditioned; used for parity).
"""
import numpy as np


def make_mixture(n_channels, n_bins, n_frames, seed=0, mode="mix", rank=4):
    rng = np.random.default_rng(seed)
    N, I, J = n_channels, n_bins, n_frames
    if mode == "iid":
        return rng.standard_normal((N, I, J)) + 1j * rng.standard_normal((N, I, J))
    if mode != "mix":
        raise ValueError("mode must be 'iid' or 'mix'")
    S = rng.standard_normal((N, I, J)) + 1j * rng.standard_normal((N, I, J))
    S = S * np.sqrt(rng.random((N, I, rank)) @ rng.random((N, rank, J)))
    A = rng.standard_normal((I, N, N)) + 1j * rng.standard_normal((I, N, N))
    return (A @ S.transpose(1, 0, 2)).transpose(1, 0, 2)


def make_batch(batch, n_channels, n_bins, n_frames, config_id=0, mode="mix"):
    """``X[B,N,I,J]`` complex128, mixture ``b`` seeded with ``1000*config_id + b``."""
    return np.stack(
        [make_mixture(n_channels, n_bins, n_frames, 1000 * config_id + b, mode) for b in range(batch)]
    )


def make_nmf_init(n_sources, n_bins, n_frames, n_basis, seed=42):
    """``T[N,I,K]`` then ``V[N,K,J]`` in the reference's draw order (ssspy/bss/ilrma.py:256-268)."""
    rng = np.random.default_rng(seed)
    T = rng.random((n_sources, n_bins, n_basis))
    V = rng.random((n_sources, n_basis, n_frames))
    return T, V
