"""Frequency-domain ICA on the device (host mirror of ssspy/bss/fdica.py: FDICABase :32-327, AuxFDICA :846-1245,
AuxLaplaceFDICA :1527-1667).

Covered: the auxiliary-function family with the Laplace contrast (``AuxLaplaceFDICA``), spatial_algorithm IP / IP1 /
IP2, permutation alignment by spectrogram correlation (``solve_permutation``), projection-back and
minimal-distortion-principle scale restoration.  Contrast functions are Python callables in the reference; on the
device they are an enum, so the generic ``AuxFDICA`` with user-defined callables and the gradient-based variants
raise ``NotImplementedError`` (no CPU fallback).
"""
import functools

import numpy as np
import torch

from .. import _device, _lib
from ..special.flooring import EPS, max_flooring
from ..utils.flooring import choose_flooring_fn
from ..utils.select_pair import sequential_pair_selector
from .base import IterativeMethodBase
from .ilrma import _not_on_device
from .iva import AuxIVA, IVABase

__all__ = ["AuxFDICA", "AuxLaplaceFDICA"]

spatial_algorithms = ["IP", "IP1", "IP2"]


class FDICABase(IVABase):
    """ssspy/bss/fdica.py:32-327.  State handling, loss plumbing and scale restoration are the IVA ones (same
    attributes: ``demix_filter``, ``output``, ``loss``); FDICA adds the permutation alignment."""

    def __init__(self, contrast_fn=None, flooring_fn=functools.partial(max_flooring, eps=EPS), callbacks=None,
                 permutation_alignment=True, scale_restoration=True, record_loss=True, reference_id=0):
        super().__init__(flooring_fn=flooring_fn, callbacks=callbacks, scale_restoration=scale_restoration,
                         record_loss=record_loss, reference_id=reference_id)
        if contrast_fn is None:
            raise ValueError("Specify contrast function.")
        self.contrast_fn = contrast_fn
        self.permutation_alignment = permutation_alignment

    def __repr__(self):
        s = "FDICA(permutation_alignment={permutation_alignment}, scale_restoration={scale_restoration}"
        s += ", record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    def solve_permutation(self):
        """fdica.py:239-255."""
        permutation_alignment = self.permutation_alignment
        assert permutation_alignment, "Set permutation_alignment=True."
        if type(permutation_alignment) is bool:
            permutation_alignment = "spectrogram_correlation"
        if permutation_alignment == "spectrogram_correlation":
            self.solve_permutation_by_correlation()
        else:
            raise NotImplementedError("permutation_alignment {} is not implemented.".format(permutation_alignment))

    def solve_permutation_by_correlation(self, flooring_fn="self"):
        """Y = W X, then ``correlation_based_permutation_solver(Y, W)`` (fdica.py:257-281): the bound ``output`` and
        ``demix_filter`` are permuted in place on the device; ``self.permutation`` keeps the chosen permutation of
        every bin, shape (n_bins, n_sources) [(batch, ...) for batched input]."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        self._set_flooring(flooring_fn)
        B, N, I, J = self._dims()
        corr = _device.empty((B, I), torch.float64)
        self._for_chunks(lambda ch, sp: _lib.call("ssb_plan_permutation_correlation", ch["plan"],
                                                  corr[ch["b0"]:].data_ptr(), sp))
        self._join()
        # host logic, as pair selectors are: numpy.argsort on float64 (permutation_alignment.py:93)
        order = torch.from_numpy(np.ascontiguousarray(np.argsort(corr.cpu().numpy(), axis=1).astype(np.int32)))
        order = order.to(corr.device)
        perms = _device.empty((B, I, N), torch.int32)
        self._for_chunks(lambda ch, sp: _lib.call("ssb_plan_permutation_align", ch["plan"], order[ch["b0"]:].data_ptr(),
                                                  perms[ch["b0"]:].data_ptr(), sp))
        self._join()
        p = perms.cpu().numpy().astype(np.int64)
        self.permutation = p if self._batched else p[0]
        self._host_output = None


class AuxFDICA(FDICABase):
    """ssspy/bss/fdica.py:846-1245 (signature :946-962)."""

    _model = None  # set by AuxLaplaceFDICA

    def __init__(self, spatial_algorithm="IP", contrast_fn=None, d_contrast_fn=None,
                 flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None, callbacks=None,
                 permutation_alignment=True, scale_restoration=True, record_loss=True, reference_id=0):
        super().__init__(contrast_fn=contrast_fn, flooring_fn=flooring_fn, callbacks=callbacks,
                         permutation_alignment=permutation_alignment, scale_restoration=scale_restoration,
                         record_loss=record_loss, reference_id=reference_id)
        assert spatial_algorithm in spatial_algorithms, "Not support {}.".format(spatial_algorithms)
        self.spatial_algorithm = spatial_algorithm
        self.d_contrast_fn = d_contrast_fn
        if pair_selector is None:
            if spatial_algorithm == "IP2":
                self.pair_selector = sequential_pair_selector
        else:
            self.pair_selector = pair_selector

    def __call__(self, input, n_iter=100, initial_call=True, **kwargs):
        """fdica.py:983-1022: iterate, align the permutations, restore the scale, separate."""
        if self._model is None:
            _not_on_device("AuxFDICA with user-defined contrast functions (use AuxLaplaceFDICA)")
        self.input = input
        self._reset(**kwargs)
        IterativeMethodBase.__call__(self, n_iter=n_iter, initial_call=initial_call)
        if self.permutation_alignment:
            self.solve_permutation()
        if self.scale_restoration:
            self.restore_scale()
        if self._state.get("demix_filter") is not None:
            self._plan_call("ssb_plan_separate")
        _lib.check_status()  # LinAlgError where the reference's np.linalg.solve / inv would have raised
        return self.output

    def __repr__(self):
        s = "AuxFDICA(spatial_algorithm={spatial_algorithm}, permutation_alignment={permutation_alignment}"
        s += ", scale_restoration={scale_restoration}, record_loss={record_loss}"
        if self.scale_restoration:
            s += ", reference_id={reference_id}"
        return (s + ")").format(**self.__dict__)

    _plan_config = AuxIVA._plan_config

    def update_once(self, flooring_fn="self"):
        """Per-bin auxiliary weights from the current separation, then IP1 / IP2 (fdica.py:1038-1245)."""
        flooring_fn = choose_flooring_fn(flooring_fn, method=self)
        if self.spatial_algorithm not in spatial_algorithms:
            raise NotImplementedError("Not support {}.".format(self.spatial_algorithm))
        self._set_flooring(flooring_fn)
        self._plan_call("ssb_update_spatial_model")

    def update_once_ip1(self, flooring_fn="self"):
        assert self.spatial_algorithm in ["IP", "IP1"]
        AuxFDICA.update_once(self, flooring_fn=flooring_fn)

    def update_once_ip2(self, flooring_fn="self"):
        assert self.spatial_algorithm == "IP2"
        AuxFDICA.update_once(self, flooring_fn=flooring_fn)


class AuxLaplaceFDICA(AuxFDICA):
    """Laplace contrast G(y) = 2 |y|, G'(y) = 2 (ssspy/bss/fdica.py:1527-1667)."""

    _model = _lib.MODEL_FDICA_LAPLACE

    def __init__(self, spatial_algorithm="IP", flooring_fn=functools.partial(max_flooring, eps=EPS), pair_selector=None,
                 callbacks=None, permutation_alignment=True, scale_restoration=True, record_loss=True, reference_id=0):
        def contrast_fn(y):
            return 2 * np.abs(y)

        def d_contrast_fn(y):
            return 2 * np.ones_like(y)

        super().__init__(spatial_algorithm=spatial_algorithm, contrast_fn=contrast_fn, d_contrast_fn=d_contrast_fn,
                         flooring_fn=flooring_fn, pair_selector=pair_selector, callbacks=callbacks,
                         permutation_alignment=permutation_alignment, scale_restoration=scale_restoration,
                         record_loss=record_loss, reference_id=reference_id)

    def __repr__(self):
        return super().__repr__().replace("AuxFDICA(", "AuxLaplaceFDICA(", 1)
