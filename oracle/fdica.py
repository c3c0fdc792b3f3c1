"""AuxLaplaceFDICA iteration and the correlation-based permutation solver (oracle; see oracle/__init__.py).

Restated from ssspy/bss/fdica.py:1065-1116 (update_once_ip1), :1118-1245 (update_once_ip2, weights recomputed
per pair), :199-237 (loss), :239-281 (permutation alignment), :283-327 (scale restoration), :983-1022 (__call__),
:1618-1653 (Laplace contrast G = 2|y|, G' = 2) and ssspy/algorithm/permutation_alignment.py:12-121.
"""
import itertools

import numpy as np

from . import spatial
from .ilrma import separate
from .projection_back import minimal_distortion_principle, projection_back


def _weight(Yabs, floor):
    """phi = G'(|y|) / floor(2 |y|) with G' = 2 (fdica.py:1104-1106, :1630-1651)."""
    return 2 * np.ones_like(Yabs) / floor(2 * Yabs)


def update_once(st, floor=spatial.max_flooring, spatial_algorithm="IP", pairs=None):
    X = st["X"]
    N = X.shape[0]
    if spatial_algorithm in ("IP", "IP1"):
        Y = separate(X, st["W"])
        phi = _weight(np.abs(Y), floor)                       # (N, I, J): per-bin weights, no cross-bin coupling
        st["W"] = spatial.update_by_ip1(st["W"], spatial.weighted_covariance(X, phi), floor)
    elif spatial_algorithm == "IP2":
        W = st["W"].copy()
        if pairs is None:
            pairs = spatial.sequential_pairs(N)
        for m, n in pairs:
            Ymn = separate(X, W[:, (m, n), :])
            phi = _weight(np.abs(Ymn), floor)
            W[:, (m, n), :] = spatial.update_by_ip2_one_pair(W, spatial.weighted_covariance(X, phi), (m, n), floor)
        st["W"] = W
    else:
        raise NotImplementedError(spatial_algorithm)


def compute_loss(st):
    """sum_i [ mean_j sum_n 2 |y| - 2 log|det W_i| ] (fdica.py:216-222)."""
    Y = separate(st["X"], st["W"])
    _, logdet = np.linalg.slogdet(st["W"])
    return float(np.sum(np.sum(np.mean(2 * np.abs(Y), axis=2), axis=0) - 2 * logdet))


def correlation_based_permutation_solver(Y, *args, floor=spatial.max_flooring):
    """ssspy/algorithm/permutation_alignment.py:12-121.  ``Y`` (n_bins, n_sources, n_frames); every array of ``args``
    (n_bins, n_sources, ...) is permuted along axis 1 like Y.  Returns (Y, args..., order, perms): the processing
    order of the bins and the permutation chosen for each bin (identity for the first one)."""
    Y = Y.copy()
    args = [a.copy() for a in args]
    n_bins, n_sources, _ = Y.shape
    permutations = list(itertools.permutations(range(n_sources)))
    P = np.abs(Y)
    norm = floor(np.sqrt(np.sum(P ** 2, axis=1, keepdims=True)))
    P = P / norm
    correlation = np.sum(P @ P.transpose(0, 2, 1), axis=(1, 2))
    indices = np.argsort(correlation)
    crit = P[indices[0]]
    perms = np.tile(np.arange(n_sources), (n_bins, 1))
    for bin_idx in range(1, n_bins):
        i = indices[bin_idx]
        best, best_perm = None, None
        for perm in permutations:
            score = np.sum(crit * P[i, perm, :])
            if best is None or score > best:
                best, best_perm = score, perm
        crit = crit + P[i, best_perm, :]
        Y[i, :] = Y[i, best_perm]
        for a in args:
            a[i, :] = a[i, best_perm]
        perms[i] = best_perm
    return (Y, *args, indices, perms)


def run(X, n_iter, W=None, floor=spatial.max_flooring, spatial_algorithm="IP", pairs=None, reference_id=0,
        permutation_alignment=True, scale_restoration=True, record_loss=True):
    """AuxFDICA.__call__ (ssspy/bss/fdica.py:983-1022 + ssspy/bss/base.py:48-77)."""
    N, I, J = X.shape
    if W is None:
        W = np.tile(np.eye(N, dtype=np.complex128), (I, 1, 1))
    st = dict(X=X.astype(np.complex128), W=W.astype(np.complex128).copy())
    loss = []
    if record_loss:
        loss.append(compute_loss(st))
    for _ in range(n_iter):
        update_once(st, floor, spatial_algorithm, pairs)
        if record_loss:
            loss.append(compute_loss(st))
    st["order"], st["perms"] = None, None
    if permutation_alignment:
        Y = separate(st["X"], st["W"]).transpose(1, 0, 2)
        Y, Wp, order, perms = correlation_based_permutation_solver(Y, st["W"], floor=floor)
        st["W"], st["order"], st["perms"] = Wp, order, perms
    if isinstance(scale_restoration, str) and scale_restoration in ("minimal_distortion_principle",
                                                                     "minimal-distortion-principle", "MDP"):
        # fdica.py:314-327: Y <- mdp(Y, X), W refitted as Y X^H (X X^H)^-1
        Y = minimal_distortion_principle(separate(st["X"], st["W"]), st["X"], reference_id)
        Xi, Yi = st["X"].transpose(1, 0, 2), Y.transpose(1, 0, 2)
        XH = np.conj(Xi.transpose(0, 2, 1))
        st["W"] = Yi @ XH @ np.linalg.inv(Xi @ XH)
    elif scale_restoration:
        st["W"] = projection_back(st["W"], reference_id=reference_id)
    st["Y"] = separate(st["X"], st["W"])
    st["loss"] = np.array(loss)
    return st
