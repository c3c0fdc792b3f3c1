"""Batched small-matrix helpers on the device (host mirror of ssspy.linalg: solve, inv2, eigh2,
eigh; ssspy/linalg/_solve.py:9-21, inv.py:4-54, eigh.py:8-207).

NumPy in -> NumPy out (complex128 / float64 on the wire, fp64 arithmetic on the device); CUDA
tensors in -> CUDA tensors out.  Matrices up to 8 x 8.
"""
import numpy as np
import torch

from .. import _device, _lib

__all__ = ["solve", "inv", "inv2", "eigh", "eigh2"]


def _prep(a):
    is_t = _device.is_tensor(a)
    real = not (a.is_complex() if is_t else np.iscomplexobj(a))
    return _device.to_device(a, torch.complex128), is_t, real


def _finish(t, is_t, real):
    if real:
        t = t.real
    return t if is_t else t.cpu().numpy()


def inv(a):
    """Batched inverse (np.linalg.inv call sites of the path: projection_back.py:89,110)."""
    A, is_t, real = _prep(a)
    N = A.shape[-1]
    assert A.shape[-2] == N, "square matrices are expected, but given shape of {}.".format(tuple(A.shape))
    out = torch.empty_like(A)
    n_mat = A.numel() // (N * N) if N else 0
    _lib.call("ssb_inv", A.data_ptr(), out.data_ptr(), n_mat, N, _device.stream_ptr())
    _lib.check_status()  # numpy.linalg.LinAlgError("Singular matrix") like np.linalg.inv
    return _finish(out, is_t, real)


def inv2(X):
    """(Multiplicative) inverse of 2x2 matrices, shape (*, 2, 2) (ssspy/linalg/inv.py:4-54)."""
    shape = tuple(X.shape)
    assert shape[-2:] == (2, 2), "2x2 matrix is expected, but given shape of {}.".format(shape)
    return inv(X)


def solve(a, b):
    """Batched ``a x = b``.  ``b`` is a stack of vectors when ``a.ndim == b.ndim + 1`` (the
    convention of ssspy/linalg/_solve.py:9-21), otherwise a stack of matrices."""
    A, is_t, real_a = _prep(a)
    Bm, _, real_b = _prep(b)
    vec = A.dim() == Bm.dim() + 1
    if vec:
        Bm = Bm.unsqueeze(-1)
    N, R = A.shape[-1], Bm.shape[-1]
    batch = torch.broadcast_shapes(A.shape[:-2], Bm.shape[:-2])
    A = A.expand(*batch, N, N).contiguous()
    Bm = Bm.expand(*batch, N, R).contiguous()
    X = torch.empty_like(Bm)
    n_mat = A.numel() // (N * N)
    _lib.call("ssb_solve", A.data_ptr(), Bm.data_ptr(), X.data_ptr(), n_mat, N, R, _device.stream_ptr())
    _lib.check_status()  # numpy.linalg.LinAlgError("Singular matrix") like np.linalg.solve (_solve.py:15)
    if vec:
        X = X[..., 0]
    return _finish(X, is_t, real_a and real_b)


def eigh(A, B=None, type=1):
    """(Generalised) Hermitian eigenproblem, ascending eigenvalues (ssspy/linalg/eigh.py:8-81).
    type 1: A z = l B z; 2: A B z = l z; 3: B A z = l z.  Eigenvector phases are the Jacobi
    solver's, not LAPACK's."""
    if type not in (1, 2, 3):
        raise ValueError("Invalid type={} is given.".format(type))
    Ad, is_t, real = _prep(A)
    N = Ad.shape[-1]
    Bd = None
    if B is not None:
        Bd, _, real_b = _prep(B)
        real = real and real_b
        Bd = Bd.expand_as(Ad).contiguous()
    lamb = torch.empty(Ad.shape[:-1], dtype=torch.float64, device=Ad.device)
    Z = torch.empty_like(Ad)
    n_mat = Ad.numel() // (N * N)
    _lib.call("ssb_eigh", Ad.data_ptr(), _device.ptr(Bd), type, lamb.data_ptr(), Z.data_ptr(), n_mat, N,
              _device.stream_ptr())
    Z = _finish(Z, is_t, real)
    return (lamb if is_t else lamb.cpu().numpy()), Z


def eigh2(A, B=None, type=1):
    """2x2 case (ssspy/linalg/eigh.py:84-161)."""
    assert tuple(A.shape[-2:]) == (2, 2), "2x2 matrix is expected, but given shape of {}.".format(tuple(A.shape))
    return eigh(A, B, type=type)
