"""Flooring-function plumbing (host mirror of ssspy/utils/flooring.py:8-24) and the mapping of the
reference's flooring callables onto the C-ABI's (mode, eps) pair."""
import functools
import inspect

from .. import _lib
from ..special.flooring import EPS, identity

_KNOWN = {"max_flooring": _lib.FLOOR_MAX, "add_flooring": _lib.FLOOR_ADD, "identity": _lib.FLOOR_NONE}


def choose_flooring_fn(flooring_fn="self", method=None):
    if flooring_fn is None:
        assert method is None, "method is given, but flooring function is not specified."
        flooring_fn = identity
    elif type(flooring_fn) is str and flooring_fn == "self":
        if method is None or not hasattr(method, "flooring_fn"):
            flooring_fn = identity
        else:
            flooring_fn = method.flooring_fn
    assert callable(flooring_fn), "flooring_fn should be callable."
    return flooring_fn


def flooring_to_enum(flooring_fn):
    """callable -> (SSB_FLOOR_*, eps).  Accepts this package's and the reference's own
    ``max_flooring`` / ``add_flooring`` / ``identity`` (optionally wrapped in functools.partial with
    ``eps``) and ``None``.  Anything else cannot run on the device: NotImplementedError
    (there is no CPU fallback)."""
    if flooring_fn is None:
        return _lib.FLOOR_NONE, 0.0
    fn, eps = flooring_fn, None
    while isinstance(fn, functools.partial):
        if "eps" in fn.keywords and eps is None:
            eps = fn.keywords["eps"]
        elif len(fn.args) >= 2 and eps is None:
            eps = fn.args[1]
        fn = fn.func
    name = getattr(fn, "__name__", None)
    if name not in _KNOWN:
        raise NotImplementedError(
            "flooring_fn {!r} cannot be mapped to a device flooring policy "
            "(supported: max_flooring, add_flooring, identity, None).".format(flooring_fn))
    if eps is None:
        try:
            eps = inspect.signature(fn).parameters["eps"].default
        except (KeyError, ValueError, TypeError):
            eps = EPS
    mode = _KNOWN[name]
    return mode, (0.0 if mode == _lib.FLOOR_NONE else float(eps))
