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
import os

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
    "latent": torch.float32,
    "diagonalizer": torch.complex64,
    "spatial": torch.float32,
}
_STATE_RANKS = {"demix_filter": 4, "output": 4, "basis": 4, "activation": 4, "variance": 3, "latent": 3,
                "diagonalizer": 4, "spatial": 4}


def no_whitening(sep):
    """``ssb_config.no_whitening``: the demixing-filter modes iterate in the whitened domain (``whitening`` attribute,
    default True; ``SSB_WHITEN=0`` switches it off process-wide for A/B runs)."""
    env = os.environ.get("SSB_WHITEN")
    if env is not None:
        return 0 if int(env) else 1
    return 0 if getattr(sep, "whitening", True) else 1


def _state_property(name):
    def getter(self):
        if name not in self._state:
            raise AttributeError(name)
        val = self._state[name]
        if name == "output" and self._host_output is not None:
            return self._host_output if self._batched else self._host_output[0]
        if _device.is_tensor(val) and val.is_cuda:
            self._join()
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


def reconstruct_nmf(self, basis, activation, latent=None):
    """``T V`` per source, or ``sum_k z_nk t_ik v_kj`` with the partitioning function (ilrma.py:297-328).
    Shapes (n_sources, n_bins, n_basis) x (n_sources, n_basis, n_frames) -> (n_sources, n_bins, n_frames), or
    (n_bins, n_basis), (n_basis, n_frames), (n_sources, n_basis) with ``latent``; leading batch axes allowed."""
    is_t = _device.is_tensor(basis)
    T = _device.to_device(basis, torch.float32)
    V = _device.to_device(activation, torch.float32)
    K = T.shape[-1]
    I, J = T.shape[-2], V.shape[-1]
    if latent is None:
        lead = tuple(T.shape[:-2])
        N = lead[-1] if lead else 1
        B = int(np.prod(lead[:-1])) if len(lead) > 1 else 1
        Z = None
    else:
        Z = _device.to_device(latent, torch.float32)
        lead = tuple(Z.shape[:-1])
        N = lead[-1]
        B = int(np.prod(lead[:-1])) if len(lead) > 1 else 1
    if V.shape[-2] != K:
        raise ValueError("basis and activation disagree on n_basis ({} vs {}).".format(K, V.shape[-2]))
    if latent is None and (T.dim() < 2 or tuple(T.shape[:-2]) != tuple(V.shape[:-2])):
        raise ValueError("basis {} and activation {} disagree on the leading axes.".format(tuple(T.shape), tuple(V.shape)))
    if latent is not None and (T.dim() != 2 or V.dim() != 2 or Z.shape[-1] != K):
        raise ValueError("with latent: basis (n_bins, n_basis), activation (n_basis, n_frames), latent (..., n_sources, "
                         "n_basis) are expected, but given {}, {}, {}.".format(tuple(T.shape), tuple(V.shape), tuple(Z.shape)))
    T, V = T.contiguous(), V.contiguous()
    R = _device.empty(lead + (I, J), torch.float32)
    _lib.call("ssb_reconstruct_nmf", T.data_ptr(), V.data_ptr(), _device.ptr(Z), R.data_ptr(), B, N, I, J, K,
              _device.stream_ptr())
    return R if is_t else R.cpu().numpy().astype(np.float64)


class DeviceSeparatorMixin:
    """Plan + buffer management.  Subclasses provide ``_plan_config()``."""

    demix_filter = _state_property("demix_filter")
    output = _state_property("output")
    basis = _state_property("basis")
    activation = _state_property("activation")
    variance = _state_property("variance")
    latent = _state_property("latent")
    diagonalizer = _state_property("diagonalizer")
    spatial = _state_property("spatial")
    # which state entries fill the (W, Y, T, V, variance) slots of ssb_plan_bind
    _plan_slots = ("demix_filter", "output", "basis", "activation", "variance")

    def _init_device_state(self):
        self._state = {}
        self._batched = False
        self._tensor_io = False
        self._cpu_tensor_io = False  # CPU (pinned) torch tensors in -> CPU tensors out
        self._dX = None
        self._input_host = None
        self._chunks = []
        self._streams = []
        self._ws_slots = []
        self._plan_key = None
        self._loss_buf = None
        self._pending_h2d = None   # host tensor whose upload is enqueued per chunk by _ensure_plan
        self._defer_ok = False     # stock __call__: upload + initial separation run at the head of each chunk's pipeline
        self._deferred_sep = None  # (W, Y) of the deferred initial separation
        self._host_output = None   # pinned host copy of `output` filled per chunk by __call__ (owned by the caller)
        self._h2d_stream = None    # copy streams of the host-tensor pipeline: uploads run back to back from t = 0 and
        self._d2h_stream = None    # downloads trail the chunks, both decoupled from the compute streams

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
        self._host_output = None
        self._pending_h2d = None
        if self._cpu_tensor_io and value.dtype == torch.complex64 and value.is_contiguous():
            # upload is deferred to the chunk streams so that it overlaps the iterations of earlier chunks
            x = _device.empty(tuple(value.shape), torch.complex64)
            self._pending_h2d = value if value.dim() == 4 else value.unsqueeze(0)
            self._input_host = None
        elif self._tensor_io:
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
        batched_rank = self._state_rank(name)
        if t.dim() == batched_rank - 1:
            t = t.unsqueeze(0)
        if name == "output":
            self._host_output = None
        cur = self._state.get(name)
        if _device.is_tensor(cur) and cur.is_cuda and cur.shape == t.shape and self._chunks:
            cur.copy_(t)  # keep the pointer the plan is bound to
        else:
            self._state[name] = t.clone() if t.data_ptr() == (value.data_ptr() if _device.is_tensor(value) else 0) else t
            self._plan_key = None  # rebinding needed

    def _state_rank(self, name):
        """Rank of the batched device layout of a state entry."""
        return _STATE_RANKS[name]

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

    # ---- plans: one ssb_plan per sub-batch ("chunk") of mixtures -----------------------------------
    # Mixtures are independent, so the batch may be cut into chunks that run to completion on their own
    # CUDA streams: host<->device copies of one chunk overlap the iterations of another, and a chunk
    # that fits the 126 MB L2 (X + P scratch) is re-read from L2 instead of HBM on every iteration.
    # ``chunk_size`` (mixtures per chunk; None = one plan over the whole batch) and ``n_streams`` are
    # plain attributes; ``SSB_CHUNK`` / ``SSB_STREAMS`` override them.
    chunk_size = None
    n_streams = 4

    def _dims(self):
        B, N, I, J = self._dX.shape
        return int(B), int(N), int(I), int(J)

    def _chunk_layout(self):
        B = self._dims()[0]
        cs = os.environ.get("SSB_CHUNK")
        cs = int(cs) if cs else self.chunk_size
        sizes = os.environ.get("SSB_HOST_LAYOUT") or getattr(self, "host_layout", None)
        if sizes and cs is None and self._cpu_tensor_io:
            # explicit chunk sizes for host tensors (experiments: the upload is the critical path, so the tail of the
            # pipeline -- iterations and download of the LAST chunk -- is what an uneven layout shortens)
            if isinstance(sizes, str):
                sizes = [int(v) for v in sizes.split(",") if v.strip()]
            out, b0 = [], 0
            for n in sizes:
                if b0 >= B:
                    break
                out.append((b0, min(b0 + int(n), B)))
                b0 += int(n)
            if b0 < B:
                out.append((b0, B))
            return out
        if cs is None and self._cpu_tensor_io and B >= 8:
            # host tensors in/out: eight chunks on four streams hide most of the PCIe copies behind the iterations
            # of the other chunks (tools/e2e_sweep.py: 22.8 ms vs 26.0 ms for one plan at config 2)
            cs = -(-B // 8)
        elif cs is None and B >= 32 and np.prod(self._dims()[1:]) <= (1 << 22):
            # device-resident input: four chunks on four streams.  Every kernel of the iteration ends in a partially
            # filled wave (1.7 - 3.7 waves per launch at config 2); the kernels of the other chunks fill those SMs
            # (measured 0.373 -> 0.345 ms per step).  Smaller, L2-sized chunks were measured slower,
            # and so is chunking when one mixture alone fills the GPU for many waves (config 4: +3 %).
            cs = -(-B // 4)
        if not cs or cs >= B:
            return [(0, B)]
        return [(b0, min(b0 + cs, B)) for b0 in range(0, B, cs)]

    def _chunk_streams(self, layout):
        """Stream of every chunk (None = the current stream when there is a single chunk)."""
        if len(layout) == 1:
            return [None]
        ns = int(os.environ.get("SSB_STREAMS", self.n_streams))
        if len(self._streams) < ns:
            self._streams = [torch.cuda.Stream() for _ in range(ns)]
        return [self._streams[ci % ns] for ci in range(len(layout))]

    def _initial_separate(self, W, Y):
        """Y = W X chunk by chunk (after the chunk's pending host->device copy of X, if any).  In a stock ``__call__``
        on a host tensor both are deferred to the head of each chunk's pipeline (``_chunk_init``): enqueued up
        front, the uploads of all chunks would sit in the streams ahead of every iteration and could not overlap
        them."""
        if self._pending_h2d is not None and self._defer_ok:
            self._deferred_sep = (W, Y)
            return
        B, N, I, J = self._dims()
        layout = self._chunk_layout()
        streams = self._chunk_streams(layout)
        cur = torch.cuda.current_stream()
        fork = None
        if len(layout) > 1:
            fork = torch.cuda.Event()
            fork.record(cur)
        for (b0, b1), st in zip(layout, streams):
            st = cur if st is None else st
            if fork is not None:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                if self._pending_h2d is not None:
                    self._dX[b0:b1].copy_(self._pending_h2d[b0:b1], non_blocking=True)
                _lib.call("ssb_separate", self._dX[b0:b1].data_ptr(), W[b0:b1].data_ptr(), Y[b0:b1].data_ptr(), b1 - b0,
                          N, I, J, st.cuda_stream)
        self._pending_h2d = None

    def _ensure_plan(self):
        """(Re)create and bind the chunk plans when shapes, options or buffers changed.  Pending
        host->device copies of the input are enqueued chunk by chunk on the chunk streams."""
        cfg = self._plan_config()
        tens = [self._dev(k) for k in self._plan_slots]
        layout = self._chunk_layout()
        key = (bytes(cfg), self._dX.data_ptr(), tuple(layout)) + tuple(_device.ptr(t) for t in tens)
        if self._chunks and key == self._plan_key:
            return
        self._destroy_plan()
        multi = len(layout) > 1
        ns = int(os.environ.get("SSB_STREAMS", self.n_streams)) if multi else 1
        streams = self._chunk_streams(layout)
        cur = torch.cuda.current_stream()
        fork = None
        if multi:
            fork = torch.cuda.Event()
            fork.record(cur)
        ws_need = 0
        for ci, (b0, b1) in enumerate(layout):
            ccfg = self._plan_config()
            ccfg.n_batch = b1 - b0
            plan = ctypes.c_void_p()
            _lib.call("ssb_plan_create", ctypes.byref(ccfg), ctypes.byref(plan))
            nbytes = ctypes.c_size_t(0)
            _lib.call("ssb_plan_workspace_bytes", plan, ctypes.byref(nbytes))
            ws_need = max(ws_need, nbytes.value, 256)
            self._chunks.append({"b0": b0, "b1": b1, "plan": plan, "slot": ci, "stream": streams[ci]})
        # One workspace per chunk (together they are as large as the workspace of a single plan over the batch).
        # A workspace holds state that outlives a call (the unweighted covariances of the power normalisation), so
        # chunks must not share one even when they share a stream.
        nch = len(self._chunks)
        if len(self._ws_slots) != nch or any(w.numel() < ws_need for w in self._ws_slots):
            self._ws_slots = [_device.empty((ws_need,), torch.uint8) for _ in range(nch)]
        for ch in self._chunks:
            b0, b1 = ch["b0"], ch["b1"]
            st = ch["stream"] if ch["stream"] is not None else cur
            if fork is not None:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                deferred = self._deferred_sep is not None
                if self._pending_h2d is not None and not deferred:
                    self._dX[b0:b1].copy_(self._pending_h2d[b0:b1], non_blocking=True)
                ws = self._ws_slots[ch["slot"]]
                ptrs = [0 if t is None else t[b0:b1].data_ptr() for t in tens]
                _lib.call("ssb_plan_bind", ch["plan"], self._dX[b0:b1].data_ptr(), ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                          ptrs[4], ws.data_ptr(), ws.numel())
                if deferred:
                    ch["init"] = True  # wait for the upload, prepare and W X at the head of this chunk's first piece of work
                else:
                    _lib.call("ssb_plan_prepare", ch["plan"], st.cuda_stream)
        if self._deferred_sep is not None:
            # every chunk's upload goes onto ONE copy stream, in chunk order, right now: the H2D engine streams the whole
            # batch at PCIe rate from the start, and a chunk's compute stream only waits for its own piece
            if self._h2d_stream is None:
                self._h2d_stream = torch.cuda.Stream()
                self._d2h_stream = torch.cuda.Stream()
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            self._h2d_stream.wait_event(ev0)
            with torch.cuda.stream(self._h2d_stream):
                for ch in self._chunks:
                    b0, b1 = ch["b0"], ch["b1"]
                    self._dX[b0:b1].copy_(self._pending_h2d[b0:b1], non_blocking=True)
                    ch["h2d_event"] = torch.cuda.Event()
                    ch["h2d_event"].record(self._h2d_stream)
        if self._deferred_sep is None:
            self._pending_h2d = None
        self._plan_key = key

    def _destroy_plan(self):
        for ch in getattr(self, "_chunks", []):
            _lib.call("ssb_plan_destroy", ch["plan"])
        self._chunks = []
        self._plan_key = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def _join(self):
        """Make the current stream wait for everything enqueued on the chunk streams."""
        if len(self._chunks) > 1:
            cur = torch.cuda.current_stream()
            for st in {id(c["stream"]): c["stream"] for c in self._chunks}.values():
                ev = torch.cuda.Event()
                ev.record(st)
                cur.wait_event(ev)

    def _for_chunks(self, fn, fork=True):
        """Enqueue ``fn(chunk, stream_ptr)`` for every chunk on its stream (after the current stream's
        work when ``fork``)."""
        self._ensure_plan()
        self._host_output = None  # any plan call may rewrite Y on the device: the cached host copy is stale
        if len(self._chunks) == 1:
            self._chunk_init(self._chunks[0], torch.cuda.current_stream())
            fn(self._chunks[0], _device.stream_ptr())
            return
        if fork:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            for st in {id(c["stream"]): c["stream"] for c in self._chunks}.values():
                st.wait_event(ev)
        for ch in self._chunks:
            with torch.cuda.stream(ch["stream"]):
                self._chunk_init(ch, ch["stream"])
                fn(ch, ch["stream"].cuda_stream)

    def _chunk_init(self, ch, st):
        """Deferred head of a chunk's pipeline: host->device copy of its mixtures, one-time plan work, Y = W X."""
        if not ch.get("init"):
            return
        ch["init"] = False
        b0, b1 = ch["b0"], ch["b1"]
        B, N, I, J = self._dims()
        st.wait_event(ch.pop("h2d_event"))  # this chunk's mixtures have arrived (copy stream, see _ensure_plan)
        _lib.call("ssb_plan_prepare", ch["plan"], st.cuda_stream)
        W, Y = self._deferred_sep
        _lib.call("ssb_separate", self._dX[b0:b1].data_ptr(), W[b0:b1].data_ptr(), Y[b0:b1].data_ptr(), b1 - b0, N, I, J,
                  st.cuda_stream)
        if not any(c.get("init") for c in self._chunks):
            self._pending_h2d = None
            self._deferred_sep = None

    def _plan_call(self, fn, *extra):
        self._for_chunks(lambda ch, sp: _lib.call(fn, ch["plan"], *extra, sp))
        self._join()

    def _set_flooring(self, flooring_fn):
        mode, eps = flooring_to_enum(flooring_fn)
        self._ensure_plan()
        for ch in self._chunks:
            _lib.call("ssb_plan_set_flooring", ch["plan"], mode, eps)

    def _loss_from_device(self):
        B = self._dims()[0]
        if self._loss_buf is None or self._loss_buf.numel() < B:
            self._loss_buf = _device.empty((B,), torch.float64)
        buf = self._loss_buf
        self._for_chunks(lambda ch, sp: _lib.call("ssb_compute_loss", ch["plan"], buf[ch["b0"]:].data_ptr(), sp))
        self._join()
        _lib.check_status()
        vals = buf[:B].cpu().numpy()
        return vals.copy() if self._batched else float(vals[0])

    def _run_iterations(self, n_iter, record_loss, initial_loss=False, tail=None):
        """n_iter x update_once for every chunk, chunk-major (each chunk runs to completion on its
        stream); ``tail(chunk, stream_ptr)`` is enqueued right behind a chunk's iterations.  Returns
        the (n_rows, B) loss array (row 0 = loss before the loop when ``initial_loss``)."""
        B = self._dims()[0]
        n_rows = (n_iter if record_loss else 0) + (1 if initial_loss else 0)
        bufs = {}

        def work(ch, sp):
            nb = ch["b1"] - ch["b0"]
            buf = None
            if n_rows:
                buf = bufs[ch["b0"]] = _device.empty((n_rows, nb), torch.float64)
            if initial_loss:
                _lib.call("ssb_compute_loss", ch["plan"], buf.data_ptr(), sp)
            if n_iter > 0:
                lp = buf[1 if initial_loss else 0:].data_ptr() if record_loss else 0
                _lib.call("ssb_run", ch["plan"], int(n_iter), lp, sp)
            if tail is not None:
                tail(ch, sp)

        self._for_chunks(work)
        self._join()
        if not n_rows:
            return None
        out = np.empty((n_rows, B))
        for b0, t in bufs.items():
            out[:, b0:b0 + t.shape[1]] = t.cpu().numpy()
        return out

    def _stock_pipeline(self, n_iter, initial_call, pb):
        """The whole stock ``__call__`` after ``_reset`` as one pipeline per chunk: [initial loss] ->
        n_iter x update_once -> scale restoration / final separate -> device->host copy of the chunk's
        output (pinned host tensors in => pinned host tensor out)."""
        self._set_flooring(self.flooring_fn)
        Y = self._dev("output")
        host = None
        if self._cpu_tensor_io:
            # a fresh pinned tensor per call: the caller owns what __call__ returned (ys = [sep(x) for x in files]
            # must not alias).  torch's caching host allocator hands a released block of the same size back without
            # a new cudaHostAlloc, so a loop that drops its results does not pay for the pinning again.
            host = torch.empty(Y.shape, dtype=Y.dtype, pin_memory=True)
        has_w = self._state.get("demix_filter") is not None

        def tail(ch, sp):
            if pb:
                _lib.call("ssb_restore_scale", ch["plan"], sp)
            elif has_w:
                _lib.call("ssb_plan_separate", ch["plan"], sp)
            if host is not None:
                # download on the copy stream: the chunk's compute stream is free for its next chunk at once
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream())
                d2h = self._d2h_stream if self._d2h_stream is not None else torch.cuda.current_stream()
                d2h.wait_event(done)
                with torch.cuda.stream(d2h):
                    host[ch["b0"]:ch["b1"]].copy_(Y[ch["b0"]:ch["b1"]], non_blocking=True)

        rec = bool(self.record_loss)
        losses = self._run_iterations(n_iter, rec, initial_loss=bool(initial_call and rec), tail=tail)
        if host is not None and self._d2h_stream is not None:
            ev = torch.cuda.Event()
            ev.record(self._d2h_stream)
            torch.cuda.current_stream().wait_event(ev)
        _lib.check_status()  # synchronises; LinAlgError where the reference's np.linalg.solve / inv would have raised
        if host is not None:
            self._host_output = host
        if losses is not None:
            self.loss.extend(losses[i].copy() if self._batched else float(losses[i, 0]) for i in range(losses.shape[0]))

    def compute_logdet(self, demix_filter):
        """``log|det W_i|`` per bin (ilrma.py:524-536, iva.py:224-236, fdica.py:225-237, mnmf.py:1263-1276)."""
        is_t = _device.is_tensor(demix_filter)
        W = _device.to_device(demix_filter, torch.complex64)
        if W.dim() < 2 or W.shape[-2] != W.shape[-1]:
            raise ValueError("square matrices (..., n_sources, n_sources) are expected, but given {}.".format(tuple(W.shape)))
        N = W.shape[-1]
        W = W.contiguous()
        out = _device.empty(tuple(W.shape[:-2]), torch.float64)
        _lib.call("ssb_logdet", W.data_ptr(), out.data_ptr(), out.numel(), N, _device.stream_ptr())
        return out if is_t else out.cpu().numpy()

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
        if X.dim() not in (3, 4) or W.dim() not in (3, 4):
            raise ValueError("input (..., n_channels, n_bins, n_frames) and demix_filter (..., n_bins, n_sources, "
                             "n_channels) are expected, but given {} and {}.".format(tuple(X.shape), tuple(W.shape)))
        Wb = W if W.dim() == 4 else W.unsqueeze(0)
        B, N, I, J = Xb.shape
        if tuple(Wb.shape[1:]) != (I, N, N) or Wb.shape[0] not in (1, B):
            # the kernel indexes W by (b * n_bins + i): a mismatching filter would be read out of bounds
            raise ValueError("demix_filter of shape {} does not match input of shape {} (determined case: "
                             "(n_bins, n_channels, n_channels) per mixture).".format(tuple(W.shape), tuple(X.shape)))
        if Wb.shape[0] != B:
            Wb = Wb.expand(B, I, N, N)  # one filter set for every mixture of the batch (NumPy broadcasting)
        Xb = Xb.contiguous()
        Y = torch.empty_like(Xb)
        _lib.call("ssb_separate", Xb.data_ptr(), Wb.contiguous().data_ptr(), Y.data_ptr(), B, N, I, J,
                  _device.stream_ptr())
        Y = Y if batched else Y[0]
        return Y if tensor_io else Y.cpu().numpy().astype(np.complex128)
