"""FastGaussMNMF iteration (oracle; see oracle/__init__.py).

Restated from SURVEY.md Appendix A.6 = ssspy/bss/mnmf.py:1278-1303 (update_once), :1305-1417 (basis,
activation), :1449-1633 (diagonaliser IP1/IP2), :1635-1675 (spatial), :632-678 (power normalisation),
:1174-1217 (separate: multichannel Wiener filter with to_psd, ssspy/special/psd.py:49-69), :1219-1261
(loss).  ``partitioning=False`` only.  State: ``X[M,I,J]``, ``T[N,I,K]``, ``V[N,K,J]``, ``Q[I,M,M]``,
``D[I,N,M]``.
"""
import numpy as np

from . import spatial


def _lamb(st):
    return st["T"] @ st["V"]  # (N, I, J)


def _L(st, lamb=None):
    """L[i,j,m] = sum_n Lambda[n,i,j] D[i,n,m] (mnmf.py:1339-1342)."""
    lamb = _lamb(st) if lamb is None else lamb
    return np.einsum("nij,inm->ijm", lamb, st["D"])


def _Z(st):
    """|Q x| with shape (I, J, M) (mnmf.py:1344-1346)."""
    return np.abs(st["Q"] @ st["X"].transpose(1, 0, 2)).transpose(0, 2, 1)


def _gh(st):
    L = _L(st)
    Z = _Z(st)
    G = np.einsum("inm,ijm->nij", st["D"], (Z / L) ** 2)   # sum_m D (Z/L)^2   (:1348-1349)
    H = np.einsum("inm,ijm->nij", st["D"], 1 / L)          # sum_m D / L       (:1350)
    return G, H


def update_basis(st, floor=spatial.max_flooring):
    G, H = _gh(st)
    num = np.einsum("nkj,nij->nik", st["V"], G)
    den = np.einsum("nkj,nij->nik", st["V"], H)
    st["T"] = floor(st["T"] * np.sqrt(num / den))


def update_activation(st, floor=spatial.max_flooring):
    G, H = _gh(st)
    num = np.einsum("nik,nij->nkj", st["T"], G)
    den = np.einsum("nik,nij->nkj", st["T"], H)
    st["V"] = floor(st["V"] * np.sqrt(num / den))


def update_diagonalizer(st, floor=spatial.max_flooring, algorithm="IP", pairs=None):
    """phi[i,m,j] = 1/L[i,j,m]; U[i,m] = mean_j phi x x^H; the same update_by_ip1/ip2 on Q
    (mnmf.py:1504-1514, :1621-1633)."""
    phi = (1 / _L(st)).transpose(2, 0, 1)  # (M, I, J)
    U = spatial.weighted_covariance(st["X"], phi)
    if algorithm in ("IP", "IP1"):
        st["Q"] = spatial.update_by_ip1(st["Q"], U, floor)
    elif algorithm == "IP2":
        st["Q"] = spatial.update_by_ip2(st["Q"], U, floor, pairs)
    else:
        raise NotImplementedError(algorithm)


def update_spatial(st):
    """D <- D sqrt( sum_j Lambda Z^2 / L^2  /  sum_j Lambda / L ), no floor (mnmf.py:1660-1675)."""
    lamb = _lamb(st)
    L = _L(st, lamb)
    Z2 = _Z(st) ** 2
    num = np.einsum("nij,ijm->inm", lamb, Z2 / L ** 2)
    den = np.einsum("nij,ijm->inm", lamb, 1 / L)
    st["D"] = np.sqrt(num / den) * st["D"]


def normalize(st, floor=spatial.max_flooring):
    """psi_m = floor(sqrt(mean_ij |Qx|_m^2)); Q[:,m,:] /= psi_m; D[:,:,m] /= psi_m^2 (mnmf.py:666-678)."""
    Z2 = _Z(st) ** 2
    psi = floor(np.sqrt(np.mean(Z2, axis=(0, 1))))
    st["Q"] = st["Q"] / psi[None, :, None]
    st["D"] = st["D"] / psi ** 2


def update_once(st, floor=spatial.max_flooring, algorithm="IP", pairs=None, normalization=True):
    update_basis(st, floor)
    update_activation(st, floor)
    update_diagonalizer(st, floor, algorithm, pairs)
    update_spatial(st)
    if normalization:
        normalize(st, floor)


def compute_loss(st):
    """sum_i( mean_j sum_m (Z^2/L + log L) - 2 log|det Q_i| ) (mnmf.py:1241-1261)."""
    L = _L(st)
    Z2 = _Z(st) ** 2
    _, logdet = np.linalg.slogdet(st["Q"])
    loss = np.sum(Z2 / L + np.log(L), axis=2)  # (I, J)
    return float((np.mean(loss, axis=-1) - 2 * logdet).sum())


def to_psd(R, floor=spatial.max_flooring):
    """ssspy/special/psd.py:49-69."""
    R = (R + np.conj(np.swapaxes(R, -2, -1))) / 2
    lam, P = np.linalg.eigh(R)
    lam = floor(lam)
    R = (P * lam[..., None, :]) @ np.conj(np.swapaxes(P, -2, -1))
    return (R + np.conj(np.swapaxes(R, -2, -1))) / 2


def separate(st, floor=spatial.max_flooring, reference_id=0):
    """Multichannel Wiener filter (mnmf.py:1186-1217): R_n = Q^-1 diag(Lambda_n D_n) Q^-H,
    R = to_psd(sum_n R_n), W^H = R^-1 R_n, y_n = (W x)[ref]."""
    lamb = _lamb(st)
    Qi = np.linalg.inv(st["Q"])  # (I, M, M)
    LD = np.einsum("nij,inm->nijm", lamb, st["D"])  # (N, I, J, M)
    Rn = np.einsum("iam,nijm,icm->nijac", Qi, LD, Qi.conj())
    R = to_psd(Rn.sum(axis=0), floor)
    WH = np.linalg.solve(R[None], Rn)  # (N, I, J, M, M)
    Wref = np.conj(WH[..., :, reference_id])  # row `ref` of W = conj of column `ref` of W^H
    return np.einsum("nijc,cij->nij", Wref, st["X"])


def run(X, T, V, Q, D, n_iter, floor=spatial.max_flooring, algorithm="IP", pairs=None, normalization=True,
        reference_id=0, record_loss=True):
    """FastGaussMNMF.__call__ (mnmf.py MNMFBase.__call__ + base.py:48-77)."""
    st = dict(X=X.astype(np.complex128), T=T.astype(np.float64).copy(), V=V.astype(np.float64).copy(),
              Q=Q.astype(np.complex128).copy(), D=D.astype(np.float64).copy())
    loss = []
    if record_loss:
        loss.append(compute_loss(st))
    for _ in range(n_iter):
        update_once(st, floor, algorithm, pairs, normalization)
        if record_loss:
            loss.append(compute_loss(st))
    st["Y"] = separate(st, floor, reference_id)
    st["loss"] = loss
    return st
