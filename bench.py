#!/usr/bin/env python
"""Benchmark of the iterative demixing hot path (BASELINE.json metric: ILRMA iterations/sec, % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]

A "step" is one ``update_once`` (source model + spatial model + normalisation) applied to the whole per-GPU batch of
mixtures.  Default workload (every N): BASELINE.json configs[1], GaussILRMA-IP, n_sources=2, n_bins=1025,
n_frames=512, n_basis=16, batch=64 mixtures PER GPU (weak scaling: the batch of independent mixtures is sharded, no
collective inside the iteration).  ``--config 4`` runs configs[3] (GaussILRMA-IP2, n_sources=8, n_bins=2049,
n_frames=1024, n_basis=32; 512 mixtures over 8 GPUs = 64 per GPU), ``--config 3`` / ``5`` configs[2] / configs[4].
One JSON line is printed by rank 0.

Timing: W >= 3 warm-up steps, then R = 5 timed regions of exactly K steps each (barrier + synchronize on both sides of
every region, CUDA events, max over ranks); the line reports the MEDIAN region (``timed_regions_ms`` lists them all).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    2: dict(tag="BASELINE configs[1]", cls="GaussILRMA", spatial="IP", n_sources=2, n_bins=1025, n_frames=512, n_basis=16,
            batch=64),
    3: dict(tag="BASELINE configs[2]", cls="AuxLaplaceIVA", spatial="ISS", n_sources=4, n_bins=1025, n_frames=512,
            n_basis=0, batch=256),
    4: dict(tag="BASELINE configs[3]: 512 mixtures sharded over 8 GPUs = 64 per GPU", cls="GaussILRMA", spatial="IP2",
            n_sources=8, n_bins=2049, n_frames=1024, n_basis=32, batch=64),
    5: dict(tag="BASELINE configs[4]", cls="FastGaussMNMF", spatial="IP", n_sources=4, n_bins=1025, n_frames=512,
            n_basis=16, batch=256),
}
REGIONS = 5


def algorithmic_bytes_per_mixture_iteration(wl):
    """X once (c64) + T,V read+write (f32) + W read+write (c64); IVA-ISS: Y read + write."""
    N, I, J, K = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"]
    if wl["cls"] == "AuxLaplaceIVA":
        return 2 * 8 * N * I * J + 2 * 4 * N * J if wl["spatial"].startswith("ISS") else 8 * N * I * J + 16 * N * N * I + 8 * N * J
    b = 8 * N * I * J + 2 * 4 * (N * I * K + N * K * J) + 2 * 8 * N * N * I
    if wl["cls"] == "FastGaussMNMF":
        b += 2 * 4 * I * N * N
    return b


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(wl):
    s = "%s-%s n_sources=%d n_bins=%d n_frames=%d" % (wl["cls"], wl["spatial"], wl["n_sources"], wl["n_bins"], wl["n_frames"])
    if wl["n_basis"]:
        s += " n_basis=%d" % wl["n_basis"]
    return s + " batch=%d per GPU (%s)" % (wl["batch"], wl["tag"])


# ---- CPU arm --------------------------------------------------------------------------------------------------------
def reference_path():
    """Directory that makes ``import ssspy`` the UNMODIFIED reference: $SSSPY_REF, the in-tree offline install
    baseline/_ref (travels to the GPU box), or /root/reference (build container only)."""
    if os.environ.get("SSB_BENCH_FORCE_PORT"):
        return None
    for p in (os.environ.get("SSSPY_REF"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isdir(os.path.join(p, "ssspy")):
            return p
    return None


def _cpu_worker(args):
    """One mixture: (warm-up) + n_iter timed ``update_once`` of the reference's own separator (kind 'reference') or of
    the oracle port (kind 'port'); returns seconds."""
    seed, n_iter, wl, ref, warm = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init
    N, I, J, K = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"]
    X = make_mixture(N, I, J, seed=seed, mode="mix")
    if ref:
        sys.path.insert(0, ref)
        import ssspy.bss.ilrma as rilrma
        import ssspy.bss.iva as riva
        import ssspy.bss.mnmf as rmnmf
        if wl["cls"] == "GaussILRMA":
            T, V = make_nmf_init(N, I, J, K)
            m = rilrma.GaussILRMA(n_basis=K, spatial_algorithm=wl["spatial"], record_loss=False, scale_restoration=False)
            m(X, n_iter=0, basis=T, activation=V)  # _reset only (ssspy/bss/ilrma.py:820-855)
        elif wl["cls"] == "AuxLaplaceIVA":
            m = riva.AuxLaplaceIVA(spatial_algorithm=wl["spatial"], record_loss=False, scale_restoration=False)
            m(X, n_iter=0)
        else:
            m = rmnmf.FastGaussMNMF(n_basis=K, record_loss=False, rng=np.random.default_rng(0))
            m(X, n_iter=0)
        step = m.update_once  # the body of the loop in ssspy/bss/base.py:68-77
    else:
        from oracle import ilrma as oilrma
        from oracle import iva as oiva
        if wl["cls"] == "GaussILRMA":
            T, V = make_nmf_init(N, I, J, K)
            st = oilrma.init_state(X, T, V, None, wl["spatial"])
            step = lambda: oilrma.update_once(st, spatial_algorithm=wl["spatial"])  # noqa: E731
        elif wl["cls"] == "AuxLaplaceIVA":
            st = oiva.init_state(X, None, wl["spatial"])
            step = lambda: oiva.update_once(st, spatial_algorithm=wl["spatial"])  # noqa: E731
        else:
            raise NotImplementedError("the oracle port of the CPU arm covers GaussILRMA and AuxLaplaceIVA")
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n_iter):
        step()
    return time.perf_counter() - t0


def cpu_arm(wl, n_iter, workers, warm=1):
    """One mixture per worker process (single-threaded BLAS) on the host cores; returns the cpu_baseline object.  The
    reference has no batch axis: mixtures are independent, so throughput = workers x n_iter / slowest worker."""
    import multiprocessing as mp
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    ref = reference_path()
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        times = pool.map(_cpu_worker, [(2000 + w, n_iter, wl, ref, warm) for w in range(workers)])
    wall = time.perf_counter() - t0
    rate = workers * n_iter / max(times)
    what = "unmodified ssspy %s.update_once (from %s)" % (wl["cls"], os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref) \
        if ref else "NumPy oracle port of the reference path"
    sample = ("%d mixtures x %d update_once after %d warm-up, one process per core, OPENBLAS_NUM_THREADS=1 (%s); "
              "slowest worker %.2f s, %.1f s wall; throughput extrapolates linearly to the batch (independent mixtures)"
              % (workers, n_iter, warm, what, max(times), wall))
    return {"value": rate, "unit": "mixture-iterations/s", "cores": workers, "kind": "reference" if ref else "port",
            "sample": sample, "numpy": np.__version__}


def cpu_arm_plan(wl, steps):
    """Bounded sample: (iterations, workers, warm-up) so that the arm ends within a few minutes.  Config 4 needs ~40 GB of
    temporaries per mixture in the reference (ssspy/bss/ilrma.py:1624-1629): few workers, one iteration, no warm-up."""
    cores = os.cpu_count() or 1
    if wl["n_sources"] >= 8 and wl["n_bins"] > 2000:
        return 1, max(1, min(cores, 4)), 0
    return max(1, min(steps, 2)), cores, 1


# ---- clocks ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML every ~20 ms on a background thread while the GPU is
    under the benchmark's load."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.th, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.rows.append((sm, reasons, util))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            time.sleep(0.02)

    def stop(self):
        self.stop_flag = True
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: %s" % self.err]}
        self.th.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(k for k, bit in names.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm)}


def synth_batch_device(B, N, I, J, seed, torch):
    """SURVEY.md 8(d) 'mix' mode on the device (only the synthetic data is made with torch ops): low-rank-variance
    sources through a random per-bin mixing matrix; complex64 (B, N, I, J)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    X = torch.empty((B, N, I, J), dtype=torch.complex64, device="cuda")
    for b in range(B):  # one mixture at a time keeps the temporaries small at config 4
        S = torch.view_as_complex(torch.randn((N, I, J, 2), generator=g, device="cuda"))
        var = torch.rand((N, I, 4), generator=g, device="cuda") @ torch.rand((N, 4, J), generator=g, device="cuda")
        S = S * torch.sqrt(var)
        A = torch.view_as_complex(torch.randn((I, N, N, 2), generator=g, device="cuda"))
        X[b] = (A @ S.permute(1, 0, 2)).permute(1, 0, 2)
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (default 2 = configs[1], the headline)")
    ap.add_argument("--batch", type=int, default=0, help="mixtures per GPU (0 = the configuration's)")
    ap.add_argument("--sources", type=int, default=0, help="n_sources (experiments: north-star N in {2, 4, 8})")
    ap.add_argument("--spatial", default="")
    ap.add_argument("--frames", type=int, default=0, help="n_frames (experiments only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--modular", action="store_true", help="disable the fused fast path")
    ap.add_argument("--chunk", type=int, default=0, help="mixtures per chunk plan (0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="chunk streams (0 = library default)")
    args = ap.parse_args()
    wl = dict(CONFIGS[args.config])
    if args.batch:
        wl["batch"] = args.batch
    if args.sources:
        wl["n_sources"] = args.sources
        wl["tag"] += ", n_sources overridden"
    if args.spatial:
        wl["spatial"] = args.spatial
    if args.frames:
        wl["n_frames"] = args.frames
    N, I, J, K, B = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"], wl["batch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 3) if args.impl == "b200" else args.warmup
    config = {"workload": workload_name(wl), "global_batch": B * max(args.gpus, 1),
              "parallelism": "batch-sharded dp%d" % args.gpus,
              "l2_policy": "inputs larger than L2 (X is %.0f MB per GPU)" % (8.0 * B * N * I * J / 1e6),
              "chunking": "chunk=%s streams=%s (0 = library default: 4 chunk plans on 4 CUDA streams for a "
                          "device-resident batch, 8 chunks for host tensors)" % (args.chunk, args.streams),
              "timed_regions": "%d regions of %d steps, median reported" % (REGIONS, steps)}

    if args.impl == "reference":
        # CPU arm: the reference's own update_once on all host cores (a bounded sample); rank 0 only.
        if rank != 0:
            return
        n_it, workers, warm = cpu_arm_plan(wl, steps)
        cb = cpu_arm(wl, n_it, workers, warm)
        rate = cb["value"]
        line = {"impl": "reference", "metric": "mixture_iterations_per_sec", "value": rate, "unit": "mixture-iterations/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * B * args.gpus / rate,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "note": "each step of this arm is a bounded sample: %d timed update_once on %d mixtures, one per host core; "
                        "ms_per_step extrapolates that rate to the %d mixtures of the GPU arm's step" % (n_it, workers, B * args.gpus),
                "e2e": {"value": rate, "unit": "mixture-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from ssspy_b200 import _lib, bss
    from ssspy_b200.utils.synth import make_nmf_init

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # synthetic shard of this rank (mixtures are independent: rank r owns [r*B, (r+1)*B)); generated on the device,
    # mirrored once into pinned host memory for the end-to-end measurement
    t_gen = time.perf_counter()
    Xd = synth_batch_device(B, N, I, J, seed=1000 * args.config + rank, torch=torch)
    torch.cuda.synchronize()
    state = {}
    if wl["cls"] == "GaussILRMA":
        T0, V0 = make_nmf_init(N, I, J, K, seed=42)
        state = dict(basis=T0, activation=V0)
    t_gen = time.perf_counter() - t_gen

    def make_sep(**kw):
        if wl["cls"] == "GaussILRMA":
            m = bss.GaussILRMA(n_basis=K, spatial_algorithm=wl["spatial"], record_loss=False, **kw)
        elif wl["cls"] == "AuxLaplaceIVA":
            m = bss.AuxLaplaceIVA(spatial_algorithm=wl["spatial"], record_loss=False, **kw)
        else:
            kw.pop("scale_restoration", None)
            m = bss.FastGaussMNMF(n_basis=K, diagonalizer_algorithm=wl["spatial"], record_loss=False,
                                  rng=np.random.default_rng(7), **kw)
        if args.modular:
            m.fast_path = False
        if args.chunk:
            m.chunk_size = args.chunk
        if args.streams:
            m.n_streams = args.streams
        return m

    def run_steps(m, k):
        if hasattr(m, "run_iterations"):
            m.run_iterations(k)  # ssb_run: K x update_once per chunk plan, chunk-major (mixtures are independent)
        else:
            for _ in range(k):
                m.update_once()

    # ---- device-resident throughput ("value") ---------------------------------------------------------------------
    def fresh_sep():
        # a new separator from the initial state (T0, V0, W = I): every timed region covers iterations W+1 .. W+K of a
        # run, the regime the reference is used in (n_iter = 100 by default), not the thousandth iteration of a state
        # that has long converged / degenerated
        m = make_sep(scale_restoration=False)
        m(Xd, n_iter=0, **state)  # binds the plans; state stays on the device
        run_steps(m, warmup)
        torch.cuda.synchronize()
        return m

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    regions, launches, sep = [], 0, None
    for r in range(REGIONS):
        del sep
        sep = fresh_sep()
        if world > 1:
            dist.barrier()
        n0 = _lib.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        run_steps(sep, steps)  # EXACTLY K steps = K x update_once over every mixture of the shard (base.py:68-77)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        launches = _lib.launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        regions.append(ms)
    ms_total = sorted(regions)[len(regions) // 2]
    if rank == 0:
        # keep the identical load running while the clock sampler collects (a timed region can be shorter than one NVML
        # sampling period); these extra steps are not timed
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            del sep
            sep = fresh_sep()
            for _ in range(4):
                run_steps(sep, steps)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the %d timed regions + 1.0 s of the same steps" % REGIONS
    if world > 1:
        dist.barrier()
    ms_per_step = ms_total / steps
    value = B * world * steps / (ms_total / 1e3)
    try:  # the state the timed iterations left behind must be a usable separator state
        _lib.check_status()
        probe = sep._dev("basis") if wl["cls"] != "AuxLaplaceIVA" else sep._dev("output")
        state_ok = bool(torch.isfinite(torch.view_as_real(probe) if probe.is_complex() else probe).all().item())
    except Exception as e:  # LinAlgError from a singular per-bin matrix
        state_ok = repr(e)
    del sep

    # ---- per-kernel times (CUDA events after every launch on the launching stream) ----------------------------------
    # The timed regions run the batch as several chunk plans on concurrent streams (engine default), where
    # event-to-event deltas of interleaved launches mean nothing; the per-kernel breakdown is therefore taken from the
    # same steps run as ONE plan on one stream through ssb_run (every launch then covers the whole per-GPU batch, which
    # is also what the algorithmic-bytes figure of the dominant kernel refers to).
    prof_steps = 5
    sep_prof = make_sep(scale_restoration=False)
    sep_prof.chunk_size = B
    sep_prof(Xd, n_iter=0, **state)
    run_steps(sep_prof, warmup)
    torch.cuda.synchronize()
    _lib.call("ssb_profile_begin", torch.cuda.current_stream().cuda_stream)
    run_steps(sep_prof, prof_steps)
    kernels = _lib.profile_end()
    del sep_prof
    total_prof = sum(k[2] for k in kernels) or 1.0
    dom = kernels[0] if kernels else ("none", 0, 0.0)
    abytes_step = algorithmic_bytes_per_mixture_iteration(wl) * B
    peak, peak_src = hbm_peak()
    achieved = abytes_step / (ms_per_step / 1e3) / 1e9
    # dominant kernel on its own: its compulsory bytes (what it must read / write even in a perfectly fused iteration)
    # over its average launch duration, and the DRAM traffic ncu measured for one launch of it (profiles/, same workload)
    x_bytes = 8 * N * I * J
    small = 4 * (N * I * K + N * K * J) if K else 0
    dom_alg = {"tma_cov_ip1_basis": x_bytes + 2 * small + 4 * N * I * J + 16 * N * N * I,
               "tma_basis": x_bytes + 2 * small + 4 * N * I * J, "coop_basis": x_bytes + 2 * small + 4 * N * I * J,
               "tma_phi_cov": x_bytes + small + 8 * N * N * N * I, "fused_phi_cov": x_bytes + small + 8 * N * N * N * I,
               "coop_phi_cov": x_bytes + small + 8 * N * N * N * I, "mma_phi_cov": x_bytes + small + 8 * N * N * N * I,
               "coop_activation": 4 * N * I * J + 2 * small}.get(dom[0])
    traffic = dram_util = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == [wl["cls"], wl["spatial"], N, I, J, K, B]:
            traffic = tj["dram_bytes_per_launch"].get(dom[0])
            step_bytes = tj.get("dram_bytes_per_step_chunked")
            if step_bytes:
                dram_util = step_bytes / (ms_per_step / 1e3) / 1e9 / peak
    except Exception:
        pass
    dom_ms = dom[2] / max(dom[1], 1)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "dram_util": dram_util, "peak_source": peak_src,
                "scope": "achieved/frac: algorithmic bytes of one whole update_once step over the per-GPU batch "
                         "(SURVEY.md 8(d): %.1f MB, %d launches) / step time; traffic: ncu DRAM bytes of one launch of the "
                         "dominant kernel; dram_util: ncu DRAM bytes of one step of the chunked run / step time / peak "
                         "(profiles/r2_ncu_traffic.json)" % (abytes_step / 1e6, launches // max(steps, 1)),
                "dominant_kernel": dom[0], "dominant_kernel_ms": dom_ms,
                "dominant_kernel_share": dom[2] / total_prof,
                "dominant_kernel_achieved": (dom_alg * B / (dom_ms / 1e3) / 1e9) if dom_alg else None,
                "dominant_kernel_frac": (dom_alg * B / (dom_ms / 1e3) / 1e9 / peak) if dom_alg else None,
                "kernels_ms_per_step": {k[0]: round(k[2] / prof_steps, 4) for k in kernels}}

    # ---- end to end through the public API with host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        X_pinned = torch.empty(Xd.shape, dtype=Xd.dtype, pin_memory=True)
        X_pinned.copy_(Xd)
        torch.cuda.synchronize()

        def e2e_once():
            m = make_sep(scale_restoration=True) if wl["cls"] != "FastGaussMNMF" else make_sep()
            t0 = time.perf_counter()
            Y = m(X_pinned, n_iter=steps, **state)  # H2D, iterate, projection back / Wiener filter, D2H
            torch.cuda.synchronize()
            return time.perf_counter() - t0, Y
        e2e_once()
        dts = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            dt, Y = e2e_once()
            if world > 1:
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            dts.append(dt)
        dt = sorted(dts)[1]
        h2d = int(X_pinned.numel() * 8 + sum(4 * v.size for v in state.values()))
        d2h = int(Y.numel() * 8)
        e2e = {"value": B * world * steps / dt, "unit": "mixture-iterations/s",
               "h2d_bytes_per_step": int(h2d / steps), "d2h_bytes_per_step": int(d2h / steps),
               "seconds_per_call": dt, "calls_s": dts,
               "pcie_floor_s": max(h2d, d2h) / 55e9,
               "note": "%s.__call__(pinned host complex64 tensor, n_iter=%d): H2D of X, whitening, %d update_once, "
                       "scale restoration, separate, D2H of Y into pinned memory; median of 3 calls; pcie_floor_s = the "
                       "larger copy at the 55 GB/s measured per direction (tools/pcie_probe.py)" % (wl["cls"], steps, steps)}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        try:
            n_it, workers, warm = cpu_arm_plan(wl, steps)
            cpu_baseline = cpu_arm(wl, n_it, workers, warm)
        except Exception as e:  # e.g. the port does not cover the configuration and the reference is not installed
            cpu_baseline = {"value": None, "unit": "mixture-iterations/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    line = {"metric": "mixture_iterations_per_sec", "value": value, "unit": "mixture-iterations/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 state, fp64 N x N solves)", "data": "synthetic",
            "config": config, "batch_iterations_per_sec": 1e3 / ms_per_step, "timed_regions_ms": regions,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "setup_seconds": {"synthetic_data": round(t_gen, 2)},
            "state_after_timed_steps_ok": state_ok}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
