"""ctypes binding of libssb.so (the C-ABI declared in include/ssb.h).

The product path has NO CPU fallback: if the library cannot be loaded, or no CUDA device is
visible when a compute entry point is reached, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libssb.so")

SSB_MAX_SOURCES = 8
SSB_MAX_BASIS = 64
SSB_MAX_PAIRS = 64

MODEL_ILRMA_GAUSS, MODEL_IVA_LAPLACE, MODEL_IVA_GAUSS, MODEL_FASTMNMF_GAUSS = 0, 1, 2, 3
MODEL_ILRMA_T, MODEL_ILRMA_GGD, MODEL_FDICA_LAPLACE = 4, 5, 6
SPATIAL_IP1, SPATIAL_IP2, SPATIAL_ISS1, SPATIAL_ISS2, SPATIAL_IPA = 0, 1, 2, 3, 4
SOURCE_MM, SOURCE_ME = 0, 1
PART_LATENT, PART_BASIS, PART_ACTIVATION = 0, 1, 2
FLOOR_MAX, FLOOR_ADD, FLOOR_NONE = 0, 1, 2
NORM_NONE, NORM_POWER, NORM_PROJECTION_BACK = 0, 1, 2


class SsbConfig(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("spatial", ctypes.c_int32), ("source", ctypes.c_int32),
        ("n_batch", ctypes.c_int32), ("n_sources", ctypes.c_int32), ("n_bins", ctypes.c_int32),
        ("n_frames", ctypes.c_int32), ("n_basis", ctypes.c_int32), ("domain", ctypes.c_float),
        ("flooring", ctypes.c_int32), ("eps", ctypes.c_float), ("normalization", ctypes.c_int32),
        ("reference_id", ctypes.c_int32), ("n_pairs", ctypes.c_int32),
        ("pairs", ctypes.c_int32 * (2 * SSB_MAX_PAIRS)), ("fast_path", ctypes.c_int32),
        ("model_param", ctypes.c_float), ("partitioning", ctypes.c_int32),
        ("ipa_normalization", ctypes.c_int32), ("ipa_newton_iter", ctypes.c_int32),
        ("no_whitening", ctypes.c_int32),
    ]


_vp, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
_i32p = ctypes.POINTER(ctypes.c_int32)

# name -> argtypes; every function returns int status (include/ssb.h)
SIGNATURES = {
    "ssb_version": [],
    "ssb_device_count": [ctypes.POINTER(ctypes.c_int)],
    "ssb_launch_count": [ctypes.POINTER(ctypes.c_ulonglong)],
    "ssb_profile_begin": [_vp],
    "ssb_profile_end": [ctypes.c_char_p, ctypes.c_size_t],
    "ssb_plan_create": [ctypes.POINTER(SsbConfig), ctypes.POINTER(_vp)],
    "ssb_plan_destroy": [_vp],
    "ssb_plan_workspace_bytes": [_vp, ctypes.POINTER(ctypes.c_size_t)],
    "ssb_plan_bind": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t],
    "ssb_plan_set_flooring": [_vp, _i, _f],
    "ssb_plan_prepare": [_vp, _vp],
    "ssb_update_once": [_vp, _vp],
    "ssb_run": [_vp, _i, _vp, _vp],
    "ssb_update_source_model": [_vp, _vp],
    "ssb_update_source_part": [_vp, _i, _vp],
    "ssb_update_spatial_model": [_vp, _vp],
    "ssb_normalize": [_vp, _vp],
    "ssb_compute_loss": [_vp, _vp, _vp],
    "ssb_restore_scale": [_vp, _vp],
    "ssb_plan_separate": [_vp, _vp],
    "ssb_restore_scale_mdp": [_vp, _vp],
    "ssb_minimal_distortion_principle": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ssb_separate": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ssb_weighted_covariance": [_vp, _vp, _ll, _ll, _ll, _i32p, _i, _vp, _i, _i, _i, _i, _vp],
    "ssb_update_by_ip1": [_vp, _vp, _i, _i, _i, _f, _vp],
    "ssb_update_by_ip2": [_vp, _vp, _i, _i, _i32p, _i, _i, _f, _vp],
    "ssb_update_by_ip2_one_pair": [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ssb_update_by_iss1": [_vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _f, _vp],
    "ssb_update_by_iss2": [_vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i32p, _i, _i, _f, _vp],
    "ssb_update_by_ipa": [_vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "ssb_permutation_correlation": [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ssb_permutation_align": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp],
    "ssb_plan_permutation_correlation": [_vp, _vp, _vp],
    "ssb_plan_permutation_align": [_vp, _vp, _vp, _vp],
    "ssb_projection_back_w": [_vp, _vp, _i, _i, _i, _vp],
    "ssb_projection_back_y": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ssb_logdet": [_vp, _vp, _i, _i, _vp],
    "ssb_reconstruct_nmf": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ssb_status_fetch": [ctypes.POINTER(ctypes.c_int), _vp],
    "ssb_inv": [_vp, _vp, _i, _i, _vp],
    "ssb_solve": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "ssb_eigh": [_vp, _vp, _i, _vp, _vp, _i, _i, _vp],
    "ssb_stft_frames": [ctypes.c_longlong, _i, _i, _vp],
    "ssb_stft": [_vp, _vp, ctypes.c_double, _vp, _i, ctypes.c_longlong, _i, _i, _vp],
    "ssb_istft": [_vp, _vp, ctypes.c_double, _vp, _vp, _i, _i, _i, _i, _vp],
    "ssb_cbrt": [_vp, _vp, ctypes.c_longlong, _vp],
    "ssb_solve_cubic": [_vp, _vp, _vp, _vp, ctypes.c_longlong, _vp],
    "ssb_lqpqm2": [_vp, _vp, _vp, _vp, _i, _i, _i, ctypes.c_double, _i, _i, _vp],
}

_lib = None


class SsbError(RuntimeError):
    pass


def load():
    """Load libssb.so; build it in-tree first if it is missing and nvcc is available."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    lib.ssb_last_error.restype = ctypes.c_char_p
    lib.ssb_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = args
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise SsbError(lib.ssb_last_error().decode("utf-8", "replace"))


STATUS_SINGULAR = 1


def check_status(stream_ptr=None):
    """Synchronise ``stream_ptr`` (the current torch stream by default) and raise what the reference would have raised
    for the work done since the last check: ``numpy.linalg.LinAlgError("Singular matrix")`` when a pivoting solve or
    inverse met an exactly zero pivot (ssspy/linalg/_solve.py:15, ssspy/algorithm/projection_back.py:89, :110)."""
    import numpy as np
    if stream_ptr is None:
        from . import _device
        stream_ptr = _device.stream_ptr()
    flags = ctypes.c_int(0)
    call("ssb_status_fetch", ctypes.byref(flags), stream_ptr)
    if flags.value & STATUS_SINGULAR:
        raise np.linalg.LinAlgError("Singular matrix")


def device_count():
    n = ctypes.c_int(0)
    load().ssb_device_count(ctypes.byref(n))
    return n.value


def launch_count():
    n = ctypes.c_ulonglong(0)
    load().ssb_launch_count(ctypes.byref(n))
    return n.value


def profile_end():
    """-> list of (kernel, launches, total_ms) sorted by total time."""
    buf = ctypes.create_string_buffer(1 << 16)
    call("ssb_profile_end", buf, len(buf))
    out = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out.append((name, int(cnt), float(ms)))
    return out


def pairs_array(pairs):
    arr = (ctypes.c_int32 * (2 * len(pairs)))()
    for q, (m, n) in enumerate(pairs):
        arr[2 * q], arr[2 * q + 1] = int(m), int(n)
    return arr
