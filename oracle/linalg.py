"""Small-matrix helpers of the hot path (oracle; see oracle/__init__.py).

Follows ssspy/linalg/_solve.py:9-21, ssspy/linalg/inv.py:4-54,
ssspy/linalg/eigh.py:8-207.  The N x N arithmetic itself lives in NumPy/LAPACK
(third party, unpinned in the reference's pyproject.toml:19-21).
"""
import numpy as np


def solve(a, b):
    """Batched ``a x = b``; ``b`` is a stack of vectors when ``a.ndim == b.ndim + 1``
    (the NumPy>=2 shim of ssspy/linalg/_solve.py:9-21)."""
    if a.ndim == b.ndim + 1:
        return np.linalg.solve(a, b[..., np.newaxis])[..., 0]
    return np.linalg.solve(a, b)


def inv2(X):
    """Closed-form 2x2 inverse: adj(X)/det(X) (ssspy/linalg/inv.py:39-54)."""
    assert X.shape[-2:] == (2, 2), "2x2 matrix is expected, but given shape of {}.".format(X.shape)
    a, b, c, d = X[..., 0, 0], X[..., 0, 1], X[..., 1, 0], X[..., 1, 1]
    det = a * d - b * c
    adj = np.stack([d, -b, -c, a], axis=-1).reshape(X.shape)
    return adj / det[..., np.newaxis, np.newaxis]


def _herm(M):
    return np.conj(np.swapaxes(M, -2, -1))


def eigh(A, B=None, type=1, _inv=np.linalg.inv):
    """(Generalised) Hermitian eigenproblem, ascending eigenvalues
    (ssspy/linalg/eigh.py:8-81 and _eigh :164-207).

    type 1: A z = l B z;  type 2: A B z = l z;  type 3: B A z = l z.
    B = L L^H (Cholesky); type 1: C = L^-1 A L^-H, z = L^-H y;
    types 2/3: C = L^H A L, z = L^-H y (2) or z = L y (3).
    """
    if B is None:
        return np.linalg.eigh(A)
    L = np.linalg.cholesky(B)
    if type == 1:
        Li = _inv(L)
        LiH = _herm(Li)
        C = Li @ A @ LiH
    elif type in (2, 3):
        LH = _herm(L)
        C = LH @ A @ L
        LiH = _inv(LH) if type == 2 else None
    else:
        raise ValueError("Invalid type={} is given.".format(type))
    lamb, y = np.linalg.eigh(C)
    z = LiH @ y if type in (1, 2) else L @ y
    return lamb, z


def eigh2(A, B=None, type=1):
    """2x2 variant: same as :func:`eigh` with the closed-form ``inv2``
    (ssspy/linalg/eigh.py:84-161)."""
    assert A.shape[-2:] == (2, 2), "2x2 matrix is expected, but given shape of {}.".format(A.shape)
    if B is None:
        return np.linalg.eigh(A)
    return eigh(A, B, type=type, _inv=inv2)


def to_psd(X, floor):
    """Hermitian-symmetrise, floor the eigenvalues, rebuild (ssspy/special/psd.py:11-71)."""
    X = (X + np.conj(np.swapaxes(X, -2, -1))) / 2
    lam, P = np.linalg.eigh(X)
    X = (P * floor(lam)[..., None, :]) @ np.conj(np.swapaxes(P, -2, -1))
    return (X + np.conj(np.swapaxes(X, -2, -1))) / 2


def psd_inv(X, floor):
    """P diag(1 / floor(lam)) P^H (ssspy/bss/_update_spatial_model.py:611-645)."""
    lam, P = np.linalg.eigh(X)
    return (P * (1 / floor(lam))[..., None, :]) @ np.conj(np.swapaxes(P, -2, -1))


def find_largest_root(A, B, C):
    """Largest real root of x^3 + A x^2 + B x + C by Cardano (ssspy/linalg/lqpqm.py:222-292)."""
    P = -A ** 2 / 3 + B
    Q = 2 * A ** 3 / 27 - A * B / 3 + C
    om = (-1 + 1j * np.sqrt(3)) / 2
    disc = ((Q / 2) ** 2 + (P / 3) ** 3).astype(np.complex128)
    w = -Q / 2 + np.sqrt(disc)
    U = np.cbrt(np.abs(w)) * np.exp(1j * np.angle(w) / 3)
    sing = U == 0
    U = np.where(sing, 1, U)
    V = -P / (3 * U)
    X1 = np.where(sing, np.cbrt(-Q), U + V)
    X2 = np.real(U * om + V * np.conj(om))
    X3 = np.real(U * np.conj(om) + V * om)
    roots = np.real(np.stack([X1, X2, X3], axis=-1))
    mono = P >= 0
    drop = mono | (~mono & (np.real(disc) > 0))  # a single real root
    roots[..., 1:] = np.where(drop[..., None], -np.inf, roots[..., 1:])
    return roots.max(axis=-1) - A / 3


def solve_equation(phi, v, z, floor, max_iter=10):
    """Largest root of f(l) = l^2 sum phi |v|^2 / (l - phi)^2 - l + z by Newton-Raphson from the cubic obtained
    with the dominant term only; coefficients normalised by phi_max (ssspy/linalg/lqpqm.py:122-219)."""
    mask = phi * np.abs(v) ** 2 >= floor(0)
    phi, v = mask * phi, mask * v
    idx = np.argmax(phi, axis=-1)
    rows = np.arange(phi.shape[0])
    phi_max = floor(phi[rows, idx])
    v_max = v[rows, idx] / phi_max
    phi, v, z = phi / phi_max[:, None], v / phi_max[:, None], z / phi_max
    A = -(np.abs(v_max) ** 2 + 2 + z)
    B = 1 + 2 * z
    C = -z
    lamb = find_largest_root(A, B, C)
    lamb = np.where(lamb > 1, lamb, 1 + floor(0))
    lamb = np.maximum(lamb, z)
    for _ in range(max_iter):
        f = lamb ** 2 * np.sum(phi * np.abs(v) ** 2 / (lamb[:, None] - phi) ** 2, axis=-1) - lamb + z
        if np.all(np.abs(f) <= floor(0)):
            break
        df = -2 * lamb * np.sum((phi * np.abs(v)) ** 2 / (lamb[:, None] - phi) ** 3, axis=-1) - 1
        mu = lamb - f / df
        lamb = np.where(mu > 1, mu, (1 + lamb) / 2)
    return lamb * phi_max


def lqpqm2(H, v, z, floor, max_iter=10):
    """argmax-side stationary point of the log-quadratically penalised quadratic minimisation, type 2
    (ssspy/linalg/lqpqm.py:13-119), singular_fn = (x < floor(0)) as update_by_ipa calls it
    (_update_spatial_model.py:484-490).  H (*, M, M) Hermitian PSD, v (*, M), z (*,).

    Known deviation (v = 0 branch only, lqpqm.py:78-89): the reference scales ``sigma_singular[:, -1]``, which is the
    last ROW of the eigenvector matrix -- one entry of every eigenvector, each with LAPACK's arbitrary phase -- where
    the derivation needs the eigenvector of the largest eigenvalue (last COLUMN).  No independent eigensolver can
    reproduce that row, so the oracle and the CUDA kernel (ssb_ipa.cu) return scale * (last column); the two agree
    with the reference in the last component and in the norm only.  The branch needs ||v|| < 1e-10, i.e. a source
    exactly uncorrelated with all others, and is not reached by the fixtures."""
    phi, sigma = np.linalg.eigh(H)
    sing = np.linalg.norm(v, axis=-1) < floor(0)
    y = np.zeros_like(v)
    if np.any(sing):
        ps, ss, zs = phi[sing], sigma[sing], z[sing]
        lam = np.maximum(zs, ps[:, -1])
        scale = np.sqrt(np.maximum((lam - zs) / ps[:, -1], 0))
        y[sing] = scale[:, None] * ss[:, :, -1]
    ns = ~sing
    if np.any(ns):
        pn, sn, vn, zn = phi[ns], sigma[ns], v[ns], z[ns]
        vt = np.sum(np.conj(sn) * vn[:, :, None], axis=-2)
        lam = solve_equation(pn, vt, zn, floor, max_iter)
        y[ns] = np.sum(sn * (pn * vt / (lam[:, None] - pn))[:, None, :], axis=-1)
    return y


def cbrt(x):
    """Cube root with complex support: cbrt(|x|) exp(i arg(x) / 3) (ssspy/linalg/cubic.py:4-22)."""
    x = np.asarray(x)
    if np.iscomplexobj(x):
        return np.cbrt(np.abs(x)) * np.exp(1j * np.angle(x) / 3)
    return np.cbrt(x)


def solve_cubic(A, B, C, D=None, all=True):
    """All three roots of A x^3 + B x^2 + C x + D (D given) or x^3 + A x^2 + B x + C, shape (3, *), Cardano in the
    reference's branch choices (ssspy/linalg/polynomial.py:9-104): principal sqrt of the discriminant, the complex cube
    root above, U := 1 and X1 := cbrt(-Q) where P == 0."""
    A, B, C = (np.asarray(t) for t in (A, B, C))
    if D is not None:
        if np.any(A == 0):
            raise np.linalg.LinAlgError("Coefficients include zero.")
        return solve_cubic(B / A, C / A, np.asarray(D) / A, all=all)
    P = (-(A ** 2) / 3 + B).astype(np.complex128)
    Q = ((2 * A ** 3) / 27 - (A * B) / 3 + C).astype(np.complex128)
    om, omc = (-1 + 1j * np.sqrt(3)) / 2, (-1 - 1j * np.sqrt(3)) / 2
    U = cbrt(-Q / 2 + np.sqrt((Q / 2) ** 2 + (P / 3) ** 3))
    sing = P == 0
    U = np.where(sing, 1, U)
    V = -P / (3 * U)
    X1 = np.where(sing, cbrt(-Q), U + V)
    X2 = np.where(sing, X1 * om, U * om + V * omc)
    X3 = np.where(sing, X1 * omc, U * omc + V * om)
    x = np.stack([X1, X2, X3], axis=0) - A / 3
    return x if all else x[0]
