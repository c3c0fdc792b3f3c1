from . import fdica, ilrma, iva, mnmf  # noqa: F401
from .fdica import AuxFDICA, AuxLaplaceFDICA  # noqa: F401
from .ilrma import GGDILRMA, TILRMA, GaussILRMA  # noqa: F401
from .iva import AuxGaussIVA, AuxIVA, AuxLaplaceIVA  # noqa: F401
from .mnmf import FastGaussMNMF  # noqa: F401
