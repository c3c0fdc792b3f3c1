"""AuxLaplaceIVA / AuxGaussIVA iteration (oracle; see oracle/__init__.py).

Restated from SURVEY.md Appendix A.4 = ssspy/bss/iva.py:1736-1793 (IP1),
:1892-1915 (IP2, weights recomputed per pair), :1958-1966 (ISS1), :3093-3115
(Laplace contrast), :3256-3289 / :3319-3337 / :3436-3473 (Gauss), :200-222 and
:2177-2192 (loss), :259-267 / :2194-2204 (projection back).
"""
import numpy as np

from . import spatial
from .ilrma import separate
from .projection_back import minimal_distortion_principle, projection_back


def init_state(X, W=None, spatial_algorithm="IP", model="laplace"):
    N, I, J = X.shape
    if W is None:
        W = np.tile(np.eye(N, dtype=np.complex128), (I, 1, 1))
    st = dict(X=X.astype(np.complex128), W=W.astype(np.complex128).copy())
    st["Y"] = separate(st["X"], st["W"])
    if spatial_algorithm in ("ISS", "ISS1", "ISS2", "IPA"):
        st["W"] = None
    if model == "gauss":
        st["variance"] = np.ones((N, J))  # ssspy/bss/iva.py:3317
    return st


def _weight(r, floor, model, variance=None):
    """phi = G'(r)/floor(2r).  Laplace G'=2 (iva.py:3105-3115); Gauss G'=2r/alpha (:3273-3289)."""
    dG = 2 * np.ones_like(r) if model == "laplace" else 2 * r / variance
    return dG / floor(2 * r)


def update_once(st, floor=spatial.max_flooring, spatial_algorithm="IP", model="laplace", pairs=None, ipa=(True, 1)):
    X = st["X"]
    N = X.shape[0]
    if model == "gauss":
        # update_source_model: alpha[n,j] = mean_i |y|^2, no floor (iva.py:3465-3473)
        Y = st["Y"] if st["W"] is None else separate(X, st["W"])
        st["variance"] = np.mean(np.abs(Y) ** 2, axis=1)
    var = st.get("variance")
    if spatial_algorithm in ("IP", "IP1"):
        Y = separate(X, st["W"])
        r = np.linalg.norm(Y, axis=1)
        phi = _weight(r, floor, model, var)
        st["W"] = spatial.update_by_ip1(st["W"], spatial.weighted_covariance(X, phi), floor)
    elif spatial_algorithm == "IP2":
        W = st["W"].copy()
        if pairs is None:
            pairs = spatial.sequential_pairs(N)
        for m, n in pairs:
            Ymn = separate(X, W[:, (m, n), :])
            r = np.linalg.norm(Ymn, axis=1)
            phi = _weight(r, floor, model, None if var is None else var[(m, n), :])
            W[:, (m, n), :] = spatial.update_by_ip2_one_pair(
                W, spatial.weighted_covariance(X, phi), (m, n), floor)
        st["W"] = W
    elif spatial_algorithm in ("ISS", "ISS1"):
        r = np.linalg.norm(st["Y"], axis=1)
        phi = _weight(r, floor, model, var)
        st["Y"] = spatial.update_by_iss1(st["Y"], phi[:, np.newaxis, :], floor)
    elif spatial_algorithm == "ISS2":  # iva.py:1968-2066: weights once, then all pairs
        r = np.linalg.norm(st["Y"], axis=1)
        phi = _weight(r, floor, model, var)
        st["Y"] = spatial.update_by_iss2(st["Y"], phi[:, np.newaxis, :], floor, pairs if pairs is not None else
                                         spatial.sequential_pairs(st["Y"].shape[0]))
    elif spatial_algorithm == "IPA":  # iva.py:2068-2176
        r = np.linalg.norm(st["Y"], axis=1)
        phi = _weight(r, floor, model, var)
        st["Y"] = spatial.update_by_ipa(st["Y"], phi[:, np.newaxis, :], floor, normalization=ipa[0], max_iter=ipa[1])
    else:
        raise NotImplementedError(spatial_algorithm)


def compute_loss(st, model="laplace"):
    """sum_n mean_j G - 2 sum_i log|det W_i| (iva.py:215-220); G = 2r (Laplace, :3103) or
    I log(alpha) + r^2/alpha (Gauss, :3267-3271); ISS: W = Y X^H (X X^H)^-1 (:2180-2187)."""
    X = st["X"]
    if st["W"] is None:
        Y = st["Y"]
        Xi, Yi = X.transpose(1, 0, 2), Y.transpose(1, 0, 2)
        XH = np.conj(Xi.transpose(0, 2, 1))
        W = Yi @ XH @ np.linalg.inv(Xi @ XH)
    else:
        W = st["W"]
        Y = separate(X, W)
    r = np.linalg.norm(Y, axis=1)
    if model == "laplace":
        G = 2 * r
    else:
        G = X.shape[1] * np.log(st["variance"]) + r ** 2 / st["variance"]
    _, logdet = np.linalg.slogdet(W)
    return float(np.sum(np.mean(G, axis=1), axis=0) - 2 * np.sum(logdet, axis=0))


def restore_scale(st, reference_id=0, method=True):
    """Projection back (ssspy/bss/iva.py:259-267, :2196-2204) or, for ``method="minimal_distortion_principle"``, the
    minimal distortion principle (:269-281, :2206-2214)."""
    if isinstance(method, str) and method in ("minimal_distortion_principle", "minimal-distortion-principle", "MDP"):
        X = st["X"]
        Y = st["Y"] if st["W"] is None else separate(X, st["W"])
        st["Y"] = minimal_distortion_principle(Y, X, reference_id)
        if st["W"] is not None:
            Xi, Yi = X.transpose(1, 0, 2), st["Y"].transpose(1, 0, 2)
            XH = np.conj(Xi.transpose(0, 2, 1))
            st["W"] = Yi @ XH @ np.linalg.inv(Xi @ XH)
        return
    if st["W"] is None:
        st["Y"] = projection_back(st["Y"], reference=st["X"], reference_id=reference_id)
    else:
        st["W"] = projection_back(st["W"], reference_id=reference_id)
        st["Y"] = separate(st["X"], st["W"])


def run(X, n_iter, W=None, floor=spatial.max_flooring, spatial_algorithm="IP", model="laplace",
        pairs=None, reference_id=0, scale_restoration=True, record_loss=True, snapshots=False, ipa=(True, 1)):
    """AuxIVA.__call__ (ssspy/bss/iva.py:1637-1672 + ssspy/bss/base.py:48-77)."""
    st = init_state(X, W, spatial_algorithm, model)
    loss, snaps = [], []
    if record_loss:
        loss.append(compute_loss(st, model))
    for _ in range(n_iter):
        update_once(st, floor, spatial_algorithm, model, pairs, ipa)
        if record_loss:
            loss.append(compute_loss(st, model))
        if snapshots:
            snaps.append({k: (None if v is None else v.copy()) for k, v in st.items() if k != "X"})
    if scale_restoration:
        restore_scale(st, reference_id, scale_restoration)
    elif st["W"] is not None:
        st["Y"] = separate(st["X"], st["W"])
    st["loss"] = loss
    st["snapshots"] = snaps
    return st
