"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from oracle import spatial as ospatial

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FLOORS = {"max": ospatial.max_flooring, "add": ospatial.add_flooring, "none": ospatial.identity}


def golden_cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def phase_align_rows(W, Wref):
    """Rotate each row of W by the unit phase that best matches Wref (IP2 eigenvector phase,
    SURVEY.md 7.3 H2)."""
    ip = np.sum(W.conj() * Wref, axis=-1, keepdims=True)
    ph = ip / np.maximum(np.abs(ip), 1e-300)
    return W * ph


def norm_arg(s):
    s = str(s)
    return {"True": True, "False": False}.get(s, s)


def sr_arg(v):
    """scale_restoration stored in a fixture: bool or keyword string."""
    v = np.asarray(v)
    return bool(v) if v.dtype == bool else str(v)


def dist_arg(g):
    """(kind, parameter) of the source distribution stored in an ILRMA fixture (Gauss when absent)."""
    kind = str(g["dist"]) if "dist" in g else "gauss"
    return (kind, float(g["dist_param"]) if kind != "gauss" else None)


def ipa_arg(g):
    """(lqpqm_normalization, newton_iter) stored in a fixture (the reference's defaults when absent)."""
    return (bool(g["ipa_normalization"]) if "ipa_normalization" in g else True,
            int(g["ipa_newton_iter"]) if "ipa_newton_iter" in g else 1)
