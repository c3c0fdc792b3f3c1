"""Per-bin demixing updates IP1 / IP2 / ISS1 (oracle; see oracle/__init__.py).

Restated from the update equations in SURVEY.md Appendix A.1(4), A.2, A.3, i.e.
ssspy/bss/_update_spatial_model.py:17-78 (IP1), :81-143 and :317-395 (IP2),
:146-194 (ISS1).  ``floor`` is a callable (max / add / identity flooring,
ssspy/special/flooring.py:6-18).
"""
import numpy as np

from .linalg import eigh2, solve


def max_flooring(x, eps=1e-10):
    return np.maximum(x, eps)


def add_flooring(x, eps=1e-10):
    return x + eps


def identity(x):
    return x


def sequential_pairs(n_sources, stop=None, step=1, sort=False):
    """ssspy/utils/select_pair.py:35-44."""
    stop = n_sources if stop is None else stop
    out = []
    for m in range(0, stop, step):
        m, n = m % n_sources, (m + 1) % n_sources
        if sort and m > n:
            m, n = n, m
        out.append((m, n))
    return out


def weighted_covariance(X, phi):
    """U[i,n,a,b] = (1/J) sum_j phi[n,i,j] X[a,i,j] conj(X[b,i,j]).

    phi is (N,I,J) (ILRMA, ssspy/bss/ilrma.py:1500-1505) or (N,J) (IVA,
    ssspy/bss/iva.py:1785-1791)."""
    N, I, J = X.shape
    Xi = X.transpose(1, 0, 2)  # (I, N, J)
    XiH = np.conj(Xi.transpose(0, 2, 1))  # (I, J, N)
    U = np.empty((I, phi.shape[0], N, N), dtype=np.complex128)
    for n in range(phi.shape[0]):
        ph = phi[n][:, np.newaxis, :] if phi.ndim == 3 else phi[n][np.newaxis, np.newaxis, :]
        U[:, n] = (Xi * ph) @ XiH
    return U / J


def update_by_ip1(W, U, floor=max_flooring):
    """For n = 0..N-1 (Gauss-Seidel, W updated in place):
    w = (W U_n)^-1 e_n;  w /= floor(sqrt(max(Re w^H U_n w, 0)));  W[:, n, :] = w^H.
    (ssspy/bss/_update_spatial_model.py:63-76)"""
    W = W.copy()
    I, N, _ = W.shape
    for n in range(N):
        Un = U[:, n]
        e = np.zeros((I, N), dtype=W.dtype)
        e[:, n] = 1
        w = solve(W @ Un, e)
        wUw = np.real(np.einsum("ia,iab,ib->i", w.conj(), Un, w))
        d = floor(np.sqrt(np.maximum(wUw, 0)))
        W[:, n, :] = w.conj() / d[:, np.newaxis]
    return W


def update_by_ip2_one_pair(W, U_pair, pair, floor=max_flooring):
    """One pairwise update (ssspy/bss/_update_spatial_model.py:353-395).

    P_q = (W U_q)^-1 [e_m e_n]  for q in (m, n)           (:356-364)
    A = P_m^H U_m P_m, B = P_n^H U_n P_n (2x2)            (:366-367)
    A h = l B h, ascending; columns flipped so h_m <- larger l   (:369-373)
    h_q /= floor(sqrt(max(Re h_q^H (A|B) h_q, 0)))        (:375-387)
    rows: W[m] = (P_m h_m)^H, W[n] = (P_n h_n)^H          (:389-393)
    Returns the (I, 2, N) pair of rows."""
    m, n = pair
    I, N, _ = W.shape
    Um, Un = U_pair[:, 0], U_pair[:, 1]
    E = np.zeros((I, N, 2), dtype=W.dtype)
    E[:, m, 0] = 1
    E[:, n, 1] = 1
    Pm = solve(W @ Um, E)
    Pn = solve(W @ Un, E)
    A = np.conj(Pm.transpose(0, 2, 1)) @ Um @ Pm
    Bm = np.conj(Pn.transpose(0, 2, 1)) @ Un @ Pn
    _, H = eigh2(A, Bm)
    H = H[..., ::-1]
    hm, hn = H[..., 0], H[..., 1]
    dm = floor(np.sqrt(np.maximum(np.real(np.einsum("ia,iab,ib->i", hm.conj(), A, hm)), 0)))
    dn = floor(np.sqrt(np.maximum(np.real(np.einsum("ia,iab,ib->i", hn.conj(), Bm, hn)), 0)))
    hm = hm / dm[:, np.newaxis]
    hn = hn / dn[:, np.newaxis]
    wm = np.einsum("iab,ib->ia", Pm, hm)
    wn = np.einsum("iab,ib->ia", Pn, hn)
    return np.stack([wm.conj(), wn.conj()], axis=1)


def update_by_ip2(W, U, floor=max_flooring, pairs=None):
    """All U computed up front, pairs in order (ssspy/bss/_update_spatial_model.py:137-141)."""
    W = W.copy()
    N = W.shape[1]
    if pairs is None:
        pairs = sequential_pairs(N)
    for m, n in pairs:
        W[:, (m, n), :] = update_by_ip2_one_pair(W, U[:, (m, n)], (m, n), floor)
    return W


def update_by_iss1(Y, phi, floor=max_flooring):
    """For n = 0..N-1 sequentially on Y (ssspy/bss/_update_spatial_model.py:181-192):
    den[m,i] = floor(mean_j phi_m |y_n|^2); v[m,i] = mean_j phi_m y_m conj(y_n) / den;
    v[n,i] = 1 - 1/sqrt(den[n,i]);  Y[m] -= v[m] y_n.
    phi broadcastable to (N,I,J)."""
    Y = Y.copy()
    N = Y.shape[0]
    for n in range(N):
        Yn = Y[n]
        num = np.mean(phi * Y * Yn.conj(), axis=-1)
        den = floor(np.mean(phi * (np.abs(Yn) ** 2), axis=-1))
        v = num / den
        v[n] = 1 - 1 / np.sqrt(den[n])
        Y = Y - v[:, :, np.newaxis] * Yn
    return Y
