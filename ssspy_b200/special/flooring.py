"""Flooring policies (host mirror of ssspy/special/flooring.py:6-18).

On the device they are an enum + eps (include/ssb.h: SSB_FLOOR_*); these callables exist so that
user code written against the reference (``functools.partial(max_flooring, eps=...)``) keeps working
and so that they can be applied to host arrays (NMF initialisation).
"""
import numpy as np

EPS = 1e-10


def identity(input):
    return input


def max_flooring(input, eps=EPS):
    return np.maximum(input, eps)


def add_flooring(input, eps=EPS):
    return input + eps
