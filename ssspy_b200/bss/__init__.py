from . import ilrma, iva  # noqa: F401
from .ilrma import GaussILRMA  # noqa: F401
from .iva import AuxGaussIVA, AuxIVA, AuxLaplaceIVA  # noqa: F401
