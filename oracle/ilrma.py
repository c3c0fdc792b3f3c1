"""GaussILRMA iteration (oracle; see oracle/__init__.py).

Restated from SURVEY.md Appendix A.1 = ssspy/bss/ilrma.py:900-922 (update_once),
:1051-1204 (MM basis/activation), :1249-1401 (ME), :1440-1696 (IP1/IP2/ISS1),
:365-514 (normalisation), :1910-1967 (loss), :538-565 / :1969-1979 (scale).
State is a plain dict: ``X[N,I,J]`` c128, ``W[I,N,N]`` c128 or None (ISS), ``Y[N,I,J]``,
``T[N,I,K]``, ``V[N,K,J]``; with the partitioning function (ilrma.py:201-245, :297-331) ``Z[N,K]``, ``T[I,K]``,
``V[K,J]`` are shared by the sources.
"""
import numpy as np

from . import spatial
from .projection_back import minimal_distortion_principle, projection_back


def separate(X, W):
    """Y[n,i,j] = sum_m W[i,n,m] X[m,i,j] (ssspy/bss/ilrma.py:292-295)."""
    return (W @ X.transpose(1, 0, 2)).transpose(1, 0, 2)


def init_state(X, T, V, W=None, spatial_algorithm="IP", Z=None):
    """ssspy/bss/ilrma.py:186-199, :897-898."""
    N, I, J = X.shape
    if W is None:
        W = np.tile(np.eye(N, dtype=np.complex128), (I, 1, 1))
    st = dict(X=X.astype(np.complex128), W=W.astype(np.complex128).copy(),
              T=T.astype(np.float64).copy(), V=V.astype(np.float64).copy())
    st["Y"] = separate(st["X"], st["W"])
    st["Z"] = None if Z is None else Z.astype(np.float64).copy()
    if spatial_algorithm in ("ISS", "ISS1", "ISS2", "IPA"):
        st["W"] = None
    return st


def reconstruct(st):
    """R[n,i,j] = sum_k T V, or sum_k z_nk t_ik v_kj with the partitioning function (ilrma.py:297-331)."""
    if st.get("Z") is None:
        return st["T"] @ st["V"]
    return np.einsum("nk,ik,kj->nij", st["Z"], st["T"], st["V"])


def update_latent(st, p=2, source_algorithm="MM", dist=("gauss", None)):
    """Z <- Z (sum_ij t v A / sum_ij t v / R)^b, then Z /= sum_n Z (ilrma.py:1007-1049; ME :1206-1247;
    t :2384-2432; GGD :3698-3743).  No flooring."""
    P = np.abs(_Y(st)) ** 2
    A, b = _source_weights(P, reconstruct(st), p, source_algorithm, dist)
    num = np.einsum("ik,kj,nij->nk", st["T"], st["V"], A)
    den = np.einsum("ik,kj,nij->nk", st["T"], st["V"], 1 / reconstruct(st))
    Z = (num / den) ** b * st["Z"]
    st["Z"] = Z / Z.sum(axis=0)


def _Y(st):
    return st["Y"] if st["W"] is None else separate(st["X"], st["W"])


def _source_weights(P, R, p, source_algorithm, dist):
    """Elementwise factors of the multiplicative update: num = sum (.) A, den = sum (.) / R, exponent b.
    Gauss MM: A = P/R^((p+2)/p), b = p/(p+2) (ilrma.py:1116-1126); Gauss ME: A = P/R^2, b = 1 (:1311-1323).
    Student-t (dof nu): R~ = nu/(nu+2) R^(2/p) + 2/(nu+2) P, A = P/(R~ R), b = p/(p+2) (MM, :2620-2640) or 1
    (ME, p = 2, :2868-2886).  GGD (beta): A = (beta/2) |y|^beta / R^((beta+p)/p), b = p/(beta+p) (:3700-3720)."""
    kind, prm = dist
    if kind == "gauss":
        return (P / R ** ((p + 2) / p), p / (p + 2)) if source_algorithm == "MM" else (P / R ** 2, 1.0)
    if kind == "t":
        nn = prm / (prm + 2)
        Rt = nn * R ** (2 / p) + (1 - nn) * P
        return P / (Rt * R), (p / (p + 2) if source_algorithm == "MM" else 1.0)
    if kind == "ggd":
        return prm / 2 * P ** (prm / 2) / R ** ((prm + p) / p), p / (prm + p)
    raise ValueError(kind)


def update_basis(st, p=2, floor=spatial.max_flooring, source_algorithm="MM", dist=("gauss", None)):
    """T <- floor(T (sum_j V A / sum_j V/R)^b)."""
    P = np.abs(_Y(st)) ** 2
    T, V = st["T"], st["V"]
    R = reconstruct(st)
    A, b = _source_weights(P, R, p, source_algorithm, dist)
    if st.get("Z") is not None:  # ilrma.py:1098-1113
        num = np.einsum("nk,kj,nij->ik", st["Z"], V, A)
        den = np.einsum("nk,kj,nij->ik", st["Z"], V, 1 / R)
    else:
        num = np.einsum("nkj,nij->nik", V, A)
        den = np.einsum("nkj,nij->nik", V, 1 / R)
    st["T"] = floor(((num / den) ** b) * T)


def update_activation(st, p=2, floor=spatial.max_flooring, source_algorithm="MM", dist=("gauss", None)):
    """Same with the new T, reduced over bins (ssspy/bss/ilrma.py:1192-1202; ME :1387-1399)."""
    P = np.abs(_Y(st)) ** 2
    T, V = st["T"], st["V"]
    R = reconstruct(st)
    A, b = _source_weights(P, R, p, source_algorithm, dist)
    if st.get("Z") is not None:  # ilrma.py:1174-1189
        num = np.einsum("nk,ik,nij->kj", st["Z"], T, A)
        den = np.einsum("nk,ik,nij->kj", st["Z"], T, 1 / R)
    else:
        num = np.einsum("nik,nij->nkj", T, A)
        den = np.einsum("nik,nij->nkj", T, 1 / R)
    st["V"] = floor(((num / den) ** b) * V)


def update_spatial(st, p=2, floor=spatial.max_flooring, spatial_algorithm="IP", pairs=None, dist=("gauss", None),
                   ipa=(True, 1)):
    """phi = 1/(T V)^(2/p) (no floor), then IP1 / IP2 on W or ISS1 on Y
    (ssspy/bss/ilrma.py:1494-1507, :1618-1633, :1690-1696).  Student-t: phi = 1/R~ (ilrma.py:2920-2934);
    GGD: phi = 1/((2/beta) floor(|y|^(2-beta)) R^(beta/p)) (:3992-4010)."""
    R = reconstruct(st)
    kind, prm = dist
    if kind == "gauss":
        phi = 1 / R ** (2 / p)
    else:
        P = np.abs(_Y(st)) ** 2
        if kind == "t":
            nn = prm / (prm + 2)
            phi = 1 / (nn * R ** (2 / p) + (1 - nn) * P)
        else:
            phi = 1 / (2 / prm * floor(P ** ((2 - prm) / 2)) * R ** (prm / p))
    if spatial_algorithm in ("IP", "IP1"):
        st["W"] = spatial.update_by_ip1(st["W"], spatial.weighted_covariance(st["X"], phi), floor)
    elif spatial_algorithm == "IP2":
        st["W"] = spatial.update_by_ip2(st["W"], spatial.weighted_covariance(st["X"], phi), floor, pairs)
    elif spatial_algorithm in ("ISS", "ISS1"):
        st["Y"] = spatial.update_by_iss1(st["Y"], phi, floor)
    elif spatial_algorithm == "ISS2":  # ilrma.py:1698-1811; class default = all sequential pairs
        st["Y"] = spatial.update_by_iss2(st["Y"], phi, floor, pairs if pairs is not None else
                                         spatial.sequential_pairs(st["Y"].shape[0]))
    elif spatial_algorithm == "IPA":  # ilrma.py:1813-1908 (lqpqm_normalization, newton_iter)
        st["Y"] = spatial.update_by_ipa(st["Y"], phi, floor, normalization=ipa[0], max_iter=ipa[1])
    else:
        raise NotImplementedError(spatial_algorithm)


def normalize(st, p=2, floor=spatial.max_flooring, normalization=True, reference_id=0):
    """Power: psi_n = floor(sqrt(mean_ij |y|^2)); T /= psi^p; W[:,n,:] /= psi (or Y)
    (ssspy/bss/ilrma.py:412-444).  Projection back: s = (W^-1)[ref,:]; W[i,n,:] *= s;
    T[n,i,:] *= |s|^p (:486-514)."""
    if normalization is True or normalization == "power":
        Y = _Y(st)
        psi = floor(np.sqrt(np.mean(np.abs(Y) ** 2, axis=(-2, -1))))
        if st.get("Z") is not None:  # ilrma.py:424-430
            Zp = st["Z"] / psi[:, None] ** p
            scale = Zp.sum(axis=0)
            st["T"] = st["T"] * scale[None, :]
            st["Z"] = Zp / scale
        else:
            st["T"] = st["T"] / psi[:, None, None] ** p
        if st["W"] is None:
            st["Y"] = Y / psi[:, None, None]
        else:
            st["W"] = st["W"] / psi[None, :, None]
    elif normalization == "projection_back":
        ref = 0 if reference_id is None else reference_id
        if st.get("Z") is not None:  # ilrma.py:466-470
            raise NotImplementedError("Projection-back-based normalization is not applicable with partitioning function.")
        if st["W"] is None:
            Y = st["Y"].transpose(1, 0, 2)
            X = st["X"].transpose(1, 0, 2)
            YH = np.conj(Y.transpose(0, 2, 1))
            scale = ((X @ YH) @ np.linalg.inv(Y @ YH))[..., ref, :]
            st["Y"] = (Y * scale[..., None]).swapaxes(-3, -2)
        else:
            scale = np.linalg.inv(st["W"])[:, ref, :]
            st["W"] = st["W"] * scale[:, :, None]
        st["T"] = st["T"] * (np.abs(scale.T) ** p)[:, :, None]
    else:
        raise NotImplementedError("Normalization {} is not implemented.".format(normalization))


def update_once(st, p=2, floor=spatial.max_flooring, spatial_algorithm="IP", source_algorithm="MM",
                normalization=True, pairs=None, reference_id=0, dist=("gauss", None), ipa=(True, 1)):
    """ssspy/bss/ilrma.py:900-922."""
    if st.get("Z") is not None:  # ilrma.py:972-973
        update_latent(st, p, source_algorithm, dist)
    update_basis(st, p, floor, source_algorithm, dist)
    update_activation(st, p, floor, source_algorithm, dist)
    update_spatial(st, p, floor, spatial_algorithm, pairs, dist, ipa)
    if normalization:
        normalize(st, p, floor, normalization, reference_id)


def compute_loss(st, p=2, dist=("gauss", None)):
    """sum_i( sum_n mean_j(P/R^(2/p) + (2/p) log TV) - 2 log|det W_i| )
    (ssspy/bss/ilrma.py:1936-1967); W-free form recovers W = Y X^H (X X^H)^-1 (:1939-1944)."""
    if st["W"] is None:
        Y = st["Y"]
        Xi, Yi = st["X"].transpose(1, 0, 2), Y.transpose(1, 0, 2)
        XH = np.conj(Xi.transpose(0, 2, 1))
        W = Yi @ XH @ np.linalg.inv(Xi @ XH)
    else:
        W = st["W"]
        Y = separate(st["X"], W)
    TV = reconstruct(st)
    kind, prm = dist
    if kind == "gauss":
        loss = np.abs(Y) ** 2 / TV ** (2 / p) + (2 / p) * np.log(TV)
    elif kind == "t":  # ilrma.py:3265-3280
        loss = (1 + prm / 2) * np.log(1 + 2 / prm * np.abs(Y) ** 2 / TV ** (2 / p)) + (2 / p) * np.log(TV)
    else:  # ilrma.py:4340-4355
        loss = np.abs(Y) ** prm / TV ** (prm / p) + (2 / p) * np.log(TV)
    _, logdet = np.linalg.slogdet(W)
    return float((np.sum(loss.mean(axis=-1), axis=0) - 2 * logdet).sum())


def restore_scale(st, reference_id=0, method=True):
    """Projection back: ssspy/bss/ilrma.py:557-565 (W-form) / :1971-1977 (Y-form); minimal distortion
    principle: :567-579 / :1981-1989."""
    if isinstance(method, str) and method in ("minimal_distortion_principle", "minimal-distortion-principle", "MDP"):
        X = st["X"]
        Y = st["Y"] if st["W"] is None else separate(X, st["W"])
        st["Y"] = minimal_distortion_principle(Y, X, reference_id)
        if st["W"] is not None:
            Xi, Yi = X.transpose(1, 0, 2), st["Y"].transpose(1, 0, 2)
            XH = np.conj(Xi.transpose(0, 2, 1))
            st["W"] = Yi @ XH @ np.linalg.inv(Xi @ XH)
            st["Y"] = separate(X, st["W"])
        return
    if st["W"] is None:
        st["Y"] = projection_back(st["Y"], reference=st["X"], reference_id=reference_id)
    else:
        st["W"] = projection_back(st["W"], reference_id=reference_id)
        st["Y"] = separate(st["X"], st["W"])


def run(X, T, V, n_iter, W=None, p=2, floor=spatial.max_flooring, spatial_algorithm="IP",
        source_algorithm="MM", normalization=True, pairs=None, reference_id=0,
        scale_restoration=True, record_loss=True, snapshots=False, dist=("gauss", None), Z=None, ipa=(True, 1)):
    """GaussILRMA.__call__ (ssspy/bss/ilrma.py:820-855 + ssspy/bss/base.py:48-77)."""
    st = init_state(X, T, V, W, spatial_algorithm, Z)
    loss, snaps = [], []
    if record_loss:
        loss.append(compute_loss(st, p, dist))
    for _ in range(n_iter):
        update_once(st, p, floor, spatial_algorithm, source_algorithm, normalization, pairs, reference_id, dist, ipa)
        if record_loss:
            loss.append(compute_loss(st, p, dist))
        if snapshots:
            snaps.append({k: (None if v is None else v.copy()) for k, v in st.items() if k != "X"})
    if scale_restoration:
        restore_scale(st, reference_id, scale_restoration)
    elif st["W"] is not None:
        st["Y"] = separate(st["X"], st["W"])
    st["loss"] = loss
    st["snapshots"] = snaps
    return st
