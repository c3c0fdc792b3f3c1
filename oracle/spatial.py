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


def update_by_iss2(Y, phi, floor=max_flooring, pairs=None):
    """Pairwise iterative source steering (ssspy/bss/_update_spatial_model.py:197-314).  For every pair (m, n), with
    u = (y_m, y_n) of the *current* Y (pair order as given, :241-261):
      other sources s: G_s = mean_j phi_s u u^H, f_s = mean_j phi_s u conj(y_s), q_s = -G_s^-1 f_s,
                       y_s <- y_s + q_s^H u                                            (:263-283)
      pair: G_m h = l G_n h (ascending, NOT flipped, :288-290); a = 0 -> (h_0, G_m), a = 1 -> (h_1, G_n);
            p_a = h_a / floor(sqrt(max(Re h_a^H G_a h_a, 0))); y_a <- p_a^H u           (:291-303)
    Default pairs: (0,1), (2,3), ... (sequential selector with step 2, :233-234).  phi broadcastable to (N,I,J)."""
    from .linalg import eigh2, inv2
    Y = Y.copy()
    N = Y.shape[0]
    phi = np.broadcast_to(phi, Y.shape)
    if pairs is None:
        pairs = sequential_pairs(N, stop=N, step=2)
    for m, n in pairs:
        m, n = m % N, n % N
        u = np.stack([Y[m], Y[n]], axis=0)                                   # (2, I, J)
        uu = (u[:, None] * u[None].conj()).transpose(2, 0, 1, 3)               # (I, 2, 2, J)
        new = Y.copy()
        for s in range(N):
            if s in (m, n):
                continue
            G = np.mean(phi[s][:, None, None, :] * uu, axis=-1)               # (I, 2, 2)
            f = np.mean(phi[s][:, None, :] * (u * Y[s].conj()[None]).transpose(1, 0, 2), axis=-1)  # (I, 2)
            q = -(inv2(G) @ f[..., None])[..., 0]
            new[s] = Y[s] + np.einsum("ic,cij->ij", q.conj(), u)
        Gm = np.mean(phi[m][:, None, None, :] * uu, axis=-1)
        Gn = np.mean(phi[n][:, None, None, :] * uu, axis=-1)
        _, H = eigh2(Gm, Gn)
        for a, (G, dst) in enumerate(((Gm, m), (Gn, n))):
            h = H[..., a]
            d = floor(np.sqrt(np.maximum(np.real(np.einsum("ia,iab,ib->i", h.conj(), G, h)), 0)))
            pa = h / d[:, None]
            new[dst] = np.einsum("ic,cij->ij", pa.conj(), u)
        Y = new
    return Y


def update_by_ipa(Y, phi, floor=max_flooring, normalization=True, max_iter=1):
    """Iterative projection with adjustment (ssspy/bss/_update_spatial_model.py:398-513).  For n = 0..N-1 on the
    current Y, per bin (s runs over the other sources, "\\n" = row / column n removed):
      U_s = to_psd(mean_j phi_s y y^H) for every s (:441-445);   Uinv = psd_inv(U_n) (:451)
      a_s = Re U_s[n,n], b_s = U_s[n,s] (:453-458);   C = conj(Uinv)\\n, d = conj(Uinv)[\\n, n] (:460-462)
      z = Re Uinv[n,n] - Re d^H C^-1 d (:464-470);   H = C / (sqrt(a_s) sqrt(a_s')), v = -b/sqrt(a) - sqrt(a) C^-1 d
      (:472-475), both divided by tr H when ``normalization`` (:477-481)
      q = lqpqm2(H, v, z) / sqrt(a) - b / a (:483-492);   qt = e_n - E conj(q);  u = U_n^-1 qt,
      p = u / floor(sqrt(max(Re qt^H u, 0))) (:494-503)
      y_n <- p^H y,  y_s <- y_s + conj(q_s) y_n(old) (:505-511).
    phi broadcastable to (N, I, J)."""
    from .linalg import lqpqm2, psd_inv, solve, to_psd
    Y = Y.copy()
    N, I, J = Y.shape
    phi = np.broadcast_to(phi, Y.shape)
    for n in range(N):
        others = [s for s in range(N) if s != n]
        YY = Y[:, None] * Y[None].conj()                                         # (a, b, I, J)
        U = np.mean(phi[:, None, None] * YY[None], axis=-1).transpose(3, 0, 1, 2)   # (I, s, a, b)
        U = to_psd(U, floor)
        Un = U[:, n]
        Uinv = psd_inv(Un, floor)
        a = np.real(U[:, others, n, n])                                           # (I, M)
        b = np.stack([U[:, s, n, s] for s in others], axis=-1)
        Cc = np.conj(Uinv)
        C = Cc[:, others][:, :, others]
        d = Cc[:, others, n]
        Cd = solve(C, d)
        z = np.real(Uinv[:, n, n]) - np.real(np.sum(d.conj() * Cd, axis=-1))
        sa = np.sqrt(a)
        H = C / (sa[:, :, None] * sa[:, None, :])
        v = -b / sa - sa * Cd
        if normalization:
            tr = np.real(np.trace(H, axis1=-2, axis2=-1))
            H = H / tr[:, None, None]
            z = z / tr
        q = lqpqm2(H, v, z, floor, max_iter) / sa - b / a
        Eq = np.zeros((I, N), dtype=np.complex128)
        Eq[:, others] = q.conj()
        qt = -Eq
        qt[:, n] += 1
        Uq = solve(Un, qt)
        den = floor(np.sqrt(np.maximum(np.real(np.sum(qt.conj() * Uq, axis=-1, keepdims=True)), 0)))
        p = Uq / den
        Yn = Y[n].copy()
        new_n = np.einsum("is,sij->ij", p.conj(), Y)
        Y = Y + Eq.T[:, :, None] * Yn[None]
        Y[n] = new_n
    return Y
