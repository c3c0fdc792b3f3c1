"""Batched small-matrix helpers on the device (host mirror of ssspy.linalg: solve, inv2, eigh2,
eigh, cbrt, solve_cubic, lqpqm2; ssspy/linalg/_solve.py:9-21, inv.py:4-54, eigh.py:8-207, cubic.py:4-22,
polynomial.py:9-104, lqpqm.py:13-119).

NumPy in -> NumPy out (complex128 / float64 on the wire, fp64 arithmetic on the device); CUDA
tensors in -> CUDA tensors out.  Matrices up to 8 x 8.
"""
import functools

import numpy as np
import torch

from .. import _device, _lib
from ..special.flooring import max_flooring
from ..utils.flooring import flooring_to_enum

EPS = 1e-10

__all__ = ["solve", "inv", "inv2", "eigh", "eigh2", "cbrt", "solve_cubic", "lqpqm2"]


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


def cbrt(x):
    """Cube root, complex values allowed: ``cbrt(|x|) exp(i arg(x) / 3)`` (ssspy/linalg/cubic.py:4-22).  Real input
    -> real output (``numpy.cbrt`` semantics: negative values keep their sign)."""
    is_t = _device.is_tensor(x)
    real = not (x.is_complex() if is_t else np.iscomplexobj(x))
    X = _device.to_device(x, torch.complex128).contiguous()
    out = torch.empty_like(X)
    _lib.call("ssb_cbrt", X.data_ptr(), out.data_ptr(), X.numel(), _device.stream_ptr())
    if real:
        # arg(x) is 0 or pi for real x: the principal complex root of a negative number is not its real cube root
        neg = X.real < 0
        out = torch.where(neg, -out.abs(), out.real)
    return out if is_t else out.cpu().numpy()


def solve_cubic(A, B, C, D=None, all=True):
    """Roots of ``A x^3 + B x^2 + C x + D = 0`` (``D`` given) or ``x^3 + A x^2 + B x + C = 0``
    (ssspy/linalg/polynomial.py:9-54): shape ``(3, *)`` complex128 in the reference's order, or the first root only
    when ``all=False``."""
    is_t = _device.is_tensor(A)
    if D is not None:
        Ad = _device.to_device(A, torch.complex128)
        if bool((Ad == 0).any()):
            raise np.linalg.LinAlgError("Coefficients include zero.")
        conv = (lambda t: t) if is_t else (lambda t: t.cpu().numpy())
        Bd, Cd, Dd = (_device.to_device(t, torch.complex128) for t in (B, C, D))
        return solve_cubic(conv(Bd / Ad), conv(Cd / Ad), conv(Dd / Ad), all=all)
    Ad, Bd, Cd = torch.broadcast_tensors(*(_device.to_device(t, torch.complex128) for t in (A, B, C)))
    Ad, Bd, Cd = Ad.contiguous(), Bd.contiguous(), Cd.contiguous()
    roots = torch.empty((3,) + tuple(Ad.shape), dtype=torch.complex128, device=Ad.device)
    _lib.call("ssb_solve_cubic", Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(), roots.data_ptr(), Ad.numel(),
              _device.stream_ptr())
    out = roots if all else roots[0]
    return out if is_t else out.cpu().numpy()


def lqpqm2(H, v, z, flooring_fn=functools.partial(max_flooring, eps=EPS), singular_fn="flooring", max_iter=10):
    """Log-quadratically penalised quadratic minimisation, type 2 (ssspy/linalg/lqpqm.py:13-119): ``H`` of shape
    (n_bins, M, M) positive semidefinite, ``v`` (n_bins, M), ``z`` (n_bins,) -> (n_bins, M).  One thread per bin: Jacobi
    eigendecomposition, Cardano start value and ``max_iter`` Newton-Raphson steps in fp64.  ``singular_fn`` may be
    "flooring" (default) or None; arbitrary callables are not supported.  In the ``v = 0`` branch the result is
    ``scale *`` (eigenvector of the largest eigenvalue), see DESIGN.md section 4 (known deviation)."""
    if singular_fn is None:
        mode = 1
    elif isinstance(singular_fn, str) and singular_fn == "flooring":
        mode = 0
    else:
        assert callable(singular_fn), "singular_fn should be callable."
        raise NotImplementedError("lqpqm2: only singular_fn='flooring' or None run on the device")
    fl_mode, eps = flooring_to_enum(flooring_fn)
    is_t = _device.is_tensor(H)
    Hd = _device.to_device(H, torch.complex128).contiguous()
    vd = _device.to_device(v, torch.complex128).contiguous()
    zd = _device.to_device(z, torch.float64).contiguous()
    M = Hd.shape[-1]
    if Hd.dim() != 3 or Hd.shape[-2] != M or tuple(vd.shape) != (Hd.shape[0], M) or tuple(zd.shape) != (Hd.shape[0],):
        raise ValueError("lqpqm2 expects H (n_bins, M, M), v (n_bins, M), z (n_bins,), but given {}, {}, {}.".format(
            tuple(Hd.shape), tuple(vd.shape), tuple(zd.shape)))
    y = torch.empty_like(vd)
    _lib.call("ssb_lqpqm2", Hd.data_ptr(), vd.data_ptr(), zd.data_ptr(), y.data_ptr(), Hd.shape[0], M, fl_mode,
              float(eps), mode, int(max_iter), _device.stream_ptr())
    return y if is_t else y.cpu().numpy()
