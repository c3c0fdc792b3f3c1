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
