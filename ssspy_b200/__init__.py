"""ssspy_b200 -- B200-native (sm_100a) iterative frequency-domain demixing behind the ssspy
separator-class API: GaussILRMA and AuxIVA (IP / IP2 / ISS), the per-bin small-matrix helpers they
use, and projection back.  Python host code -> ctypes -> libssb.so (hand-written CUDA); PyTorch only
provides device buffers and streams.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import algorithm, bss, io, linalg, special, utils  # noqa: F401,E402
from .io import wavread, wavwrite  # noqa: F401,E402
