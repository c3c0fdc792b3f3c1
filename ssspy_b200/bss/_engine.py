"""Device-resident separator state shared by GaussILRMA and the AuxIVA family.

State lives on the GPU as PyTorch tensors (device buffers only; no torch op computes anything on
the demixing path) bound to an ``ssb_plan`` of libssb.so.  The reference's attributes (``input``,
``output``, ``demix_filter``, ``basis``, ``activation``, ``variance``) are exposed as properties
that materialise NumPy arrays lazily on read and re-upload on write, so callbacks that inspect or
mutate the separator every iteration (ssspy/bss/base.py:75-77) keep working.

Extension over the reference: ``input`` may be ``(B, N, I, J)`` -- a batch of independent
mixtures -- and may be a CUDA ``torch.Tensor`` (zero-copy in, tensors out).
"""
import ctypes

import numpy as np
import torch

from .. import _device, _lib
from ..utils.flooring import flooring_to_enum

_STATE_DTYPES = {
    "demix_filter": torch.complex64,
    "output": torch.complex64,
    "basis": torch.float32,
    "activation": torch.float32,
    "variance": torch.float32,
}


def _state_property(name):
    def getter(self):
        if name not in self._state:
            raise AttributeError(name)
        val = self._state[name]
        if val is None or not _device.is_tensor(val) or self._tensor_io:
            if _device.is_tensor(val) and self._cpu_tensor_io:
                host = torch.empty(val.shape, dtype=val.dtype, pin_memory=True)
                host.copy_(val)  # device -> pinned host
                val = host
            if _device.is_tensor(val) and not self._batched:
                return val[0]
            return val
        arr = val.detach().cpu().numpy()
        arr = arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)
        return arr if self._batched else arr[0]

    def setter(self, value):
        self._set_state(name, value)

    def deleter(self):
        self._state.pop(name, None)

    return property(getter, setter, deleter)


class DeviceSeparatorMixin:
    """Plan + buffer management.  Subclasses provide ``_plan_config()``."""

    demix_filter = _state_property("demix_filter")
    output = _state_property("output")
    basis = _state_property("basis")
    activation = _state_property("activation")
    variance = _state_property("variance")

    def _init_device_state(self):
        self._state = {}
        self._batched = False
        self._tensor_io = False
        self._cpu_tensor_io = False  # CPU (pinned) torch tensors in -> CPU tensors out
        self._dX = None
        self._input_host = None
        self._plan = None
        self._plan_key = None
        self._ws = None
        self._loss_buf = None

    # ---- input -----------------------------------------------------------------------------------
    @property
    def input(self):
        if self._dX is None:
            return None
        if self._tensor_io:
            return self._dX if self._batched else self._dX[0]
        return self._input_host

    @input.setter
    def input(self, value):
        if value is None:
            self._dX = None
            self._input_host = None
            return
        self._tensor_io = _device.is_tensor(value)
        self._cpu_tensor_io = self._tensor_io and not value.is_cuda
        if value.ndim not in (3, 4):
            raise ValueError("input must have shape (n_channels, n_bins, n_frames) or "
                             "(batch, n_channels, n_bins, n_frames), but given {}.".format(tuple(value.shape)))
        self._batched = value.ndim == 4
        if self._tensor_io:
            x = _device.to_device(value, torch.complex64)
            self._input_host = None
        else:
            self._input_host = np.array(value)  # the reference keeps a private copy (ilrma.py:840)
            x = _device.to_device(self._input_host, torch.complex64)
        # CUDA tensors are bound zero-copy: the kernels only read X
        self._dX = x if self._batched else x.unsqueeze(0)

    # ---- generic state ---------------------------------------------------------------------------
    def _set_state(self, name, value):
        if value is None:
            self._state[name] = None
            return
        if self._dX is None:
            # dimensions unknown yet: keep the host value, uploaded by _reset
            self._state[name] = value if _device.is_tensor(value) else np.array(value)
            return
        t = _device.to_device(value, _STATE_DTYPES[name])
        batched_rank = {"demix_filter": 4, "output": 4, "basis": 4, "activation": 4, "variance": 3}[name]
        if t.dim() == batched_rank - 1:
            t = t.unsqueeze(0)
        cur = self._state.get(name)
        if _device.is_tensor(cur) and cur.is_cuda and cur.shape == t.shape and self._plan is not None:
            cur.copy_(t)  # keep the pointer the plan is bound to
        else:
            self._state[name] = t.clone() if t.data_ptr() == (value.data_ptr() if _device.is_tensor(value) else 0) else t
            self._plan_key = None  # rebinding needed

    def _dev(self, name):
        """Device tensor of a state entry (uploading a pending host value first)."""
        val = self._state.get(name)
        if val is None:
            return None
        if not (_device.is_tensor(val) and val.is_cuda and val.dtype == _STATE_DTYPES[name]):
            del self._state[name]
            self._set_state(name, val)
            val = self._state[name]
        return val

    def _has(self, name):
        return name in self._state

    # ---- plan ------------------------------------------------------------------------------------
    def _dims(self):
        B, N, I, J = self._dX.shape
        return int(B), int(N), int(I), int(J)

    def _ensure_plan(self):
        """(Re)create and bind the ssb_plan when shapes, options or buffers changed."""
        cfg = self._plan_config()
        ptrs = tuple(_device.ptr(self._dev(k)) for k in ("demix_filter", "output", "basis", "activation", "variance"))
        key = (bytes(cfg), self._dX.data_ptr()) + ptrs
        if self._plan is not None and key == self._plan_key:
            return
        self._destroy_plan()
        plan = ctypes.c_void_p()
        _lib.call("ssb_plan_create", ctypes.byref(cfg), ctypes.byref(plan))
        self._plan = plan
        nbytes = ctypes.c_size_t(0)
        _lib.call("ssb_plan_workspace_bytes", plan, ctypes.byref(nbytes))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = _device.empty((max(nbytes.value, 256),), torch.uint8)
        _lib.call("ssb_plan_bind", plan, self._dX.data_ptr(), ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4],
                  self._ws.data_ptr(), self._ws.numel())
        _lib.call("ssb_plan_prepare", plan, _device.stream_ptr())
        self._plan_key = key

    def _destroy_plan(self):
        if getattr(self, "_plan", None) is not None:
            _lib.call("ssb_plan_destroy", self._plan)
            self._plan = None
            self._plan_key = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def _plan_call(self, fn, *extra):
        self._ensure_plan()
        _lib.call(fn, self._plan, *extra, _device.stream_ptr())

    def _set_flooring(self, flooring_fn):
        mode, eps = flooring_to_enum(flooring_fn)
        self._ensure_plan()
        _lib.call("ssb_plan_set_flooring", self._plan, mode, eps)

    def _loss_from_device(self):
        B = self._dims()[0]
        if self._loss_buf is None or self._loss_buf.numel() < B:
            self._loss_buf = _device.empty((B,), torch.float64)
        self._plan_call("ssb_compute_loss", self._loss_buf.data_ptr())
        vals = self._loss_buf[:B].cpu().numpy()
        return vals.copy() if self._batched else float(vals[0])

    # ---- separate --------------------------------------------------------------------------------
    def separate(self, input, demix_filter):
        """``Y = W X`` per bin (ssspy/bss/ilrma.py:272-295, ssspy/bss/iva.py:171-194)."""
        if demix_filter is None:
            # the reference fails the same way on ``None @ ndarray``
            raise ValueError("matmul: Input operand 0 does not have enough dimensions "
                             "(demix_filter is None; ISS-mode separators keep their state in `output`)")
        tensor_io = _device.is_tensor(input)
        X = _device.to_device(input, torch.complex64)
        W = _device.to_device(demix_filter, torch.complex64)
        batched = X.dim() == 4
        Xb = X if batched else X.unsqueeze(0)
        Wb = W if W.dim() == 4 else W.unsqueeze(0)
        B, N, I, J = Xb.shape
        Y = torch.empty_like(Xb)
        _lib.call("ssb_separate", Xb.data_ptr(), Wb.contiguous().data_ptr(), Y.data_ptr(), B, N, I, J,
                  _device.stream_ptr())
        Y = Y if batched else Y[0]
        return Y if tensor_io else Y.cpu().numpy().astype(np.complex128)
