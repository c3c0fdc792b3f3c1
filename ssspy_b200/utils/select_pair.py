"""Pair schedules for IP2 (host mirror of ssspy/utils/select_pair.py:5-76).

The schedule is evaluated on the host and shipped to the kernels as an int32 list, so it is
bit-exact by construction; ``wrap_pairs`` applies NumPy's negative-index wrapping
(ssspy/bss/_update_spatial_model.py:241-244, tests/package/bss/test_update_spatial_model.py:19-24).
"""
import itertools


def sequential_pair_selector(n_sources, stop=None, step=1, sort=False):
    """Yield (m, (m+1) mod N) for m = 0, step, 2*step, ... < stop (default stop = N)."""
    last = n_sources if stop is None else stop
    for start in range(0, last, step):
        a, b = start % n_sources, (start + 1) % n_sources
        if sort and a > b:
            a, b = b, a
        yield a, b


def combination_pair_selector(n_sources, sort=False):
    """Yield every 2-combination of range(N) in lexicographic order."""
    for a, b in itertools.combinations(range(n_sources), 2):
        if sort and a > b:
            a, b = b, a
        yield a, b


def wrap_pairs(pairs, n_sources):
    out = []
    for m, n in pairs:
        m, n = int(m), int(n)
        if not (-n_sources <= m < n_sources and -n_sources <= n < n_sources):
            raise IndexError("pair ({}, {}) is out of bounds for {} sources".format(m, n, n_sources))
        out.append((m % n_sources, n % n_sources))
    return out


def wrap_reference_id(reference_id, n_channels):
    """``reference_id`` as an index into ``[0, n_channels)``: negative values wrap as NumPy indexing does in the
    reference (``projection_back.py:94``, ``x[reference_id]``); out-of-range values raise ``IndexError`` like it."""
    ref = int(reference_id)
    if not -n_channels <= ref < n_channels:
        raise IndexError("index {} is out of bounds for axis 0 with size {}".format(ref, n_channels))
    return ref % n_channels
