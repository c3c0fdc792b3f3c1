#!/usr/bin/env python
"""Benchmark of the ILRMA hot path (BASELINE.json metric: ILRMA iterations/sec, % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one ``update_once`` (source model + spatial model + normalisation) applied to the
whole per-GPU batch of mixtures.  Workload at every N: BASELINE.json configs[1], GaussILRMA-IP,
n_sources=2, n_bins=1025, n_frames=512, n_basis=16, batch=64 mixtures PER GPU (weak scaling: the
batch of independent mixtures is sharded, no collective inside the iteration).  One JSON line is
printed by rank 0.
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

WORKLOAD = dict(model="GaussILRMA", spatial="IP", n_sources=2, n_bins=1025, n_frames=512, n_basis=16, batch=64)


def algorithmic_bytes_per_mixture_iteration(N, I, J, K):
    """X once (c64) + T,V read+write (f32) + W read+write (c64)."""
    return 8 * N * I * J + 2 * 4 * (N * I * K + N * K * J) + 2 * 8 * N * N * I


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU arm: the oracle (NumPy restatement of the reference path) on the host cores ---------------
def _cpu_worker(args):
    seed, n_iter, wl = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ilrma as oilrma
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init
    N, I, J, K = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"]
    X = make_mixture(N, I, J, seed=seed, mode="mix")
    T, V = make_nmf_init(N, I, J, K)
    st = oilrma.init_state(X, T, V, None, wl["spatial"])
    oilrma.update_once(st, spatial_algorithm=wl["spatial"])  # warm-up
    t0 = time.perf_counter()
    for _ in range(n_iter):
        oilrma.update_once(st, spatial_algorithm=wl["spatial"])
    return time.perf_counter() - t0


def cpu_arm(wl, n_iter, workers):
    """One mixture per worker process (single-threaded BLAS), all host cores; returns
    (mixture-iterations/s, cores used, description of the sample)."""
    import multiprocessing as mp
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        times = pool.map(_cpu_worker, [(2000 + w, n_iter, wl) for w in range(workers)])
    wall = time.perf_counter() - t0
    rate = workers * n_iter / max(times)  # slowest worker bounds the parallel throughput
    sample = "%d mixtures x %d update_once, one process per core (NumPy oracle, OPENBLAS_NUM_THREADS=1), %.1f s wall" % (
        workers, n_iter, wall)
    return rate, workers, sample


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML every ~20 ms on a background thread
    while the GPU is under the benchmark's load."""

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch"], help="mixtures per GPU")
    ap.add_argument("--sources", type=int, default=WORKLOAD["n_sources"])
    ap.add_argument("--spatial", default=WORKLOAD["spatial"])
    ap.add_argument("--frames", type=int, default=WORKLOAD["n_frames"], help="n_frames (experiments only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--modular", action="store_true", help="disable the fused fast path")
    ap.add_argument("--chunk", type=int, default=0, help="mixtures per chunk plan (0 = library default)")
    ap.add_argument("--streams", type=int, default=0, help="chunk streams (0 = library default)")
    args = ap.parse_args()
    wl = dict(WORKLOAD, batch=args.batch, n_sources=args.sources, spatial=args.spatial, n_frames=args.frames)
    N, I, J, K, B = wl["n_sources"], wl["n_bins"], wl["n_frames"], wl["n_basis"], wl["batch"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = args.steps, max(args.warmup, 3) if args.impl == "b200" else args.warmup
    config = {"workload": "GaussILRMA-%s n_sources=%d n_bins=%d n_frames=%d n_basis=%d batch=%d per GPU (BASELINE configs[1])"
              % (wl["spatial"], N, I, J, K, B), "global_batch": B * max(args.gpus, 1), "parallelism": "batch-sharded dp%d" % args.gpus,
              "l2_policy": "inputs larger than L2 (X is %.0f MB per GPU)" % (8.0 * B * N * I * J / 1e6),
              "chunking": "chunk=%s streams=%s (0 = library default: 4 chunk plans on 4 CUDA streams for a "
                          "device-resident batch, 8 chunks for host tensors)" % (args.chunk, args.streams)}

    if args.impl == "reference":
        # CPU arm: the reference path's NumPy restatement on all host cores; rank 0 only.
        if rank != 0:
            return
        workers = os.cpu_count() or 1
        n_it = max(1, min(steps, 4))
        # (each worker does its own untimed warm-up iteration, see _cpu_worker)
        rate, cores, sample = cpu_arm(wl, n_it, workers)
        line = {"impl": "reference", "metric": "mixture_iterations_per_sec", "value": rate, "unit": "mixture-iterations/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * B * args.gpus / rate,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": {"value": rate, "unit": "mixture-iterations/s", "cores": cores,
                                                   "kind": "port", "sample": sample},
                "e2e": {"value": rate, "unit": "mixture-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from ssspy_b200 import _lib
    from ssspy_b200.bss import GaussILRMA
    from ssspy_b200.utils.synth import make_mixture, make_nmf_init

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # synthetic batch of this rank (mixtures are independent: rank r owns [r*B, (r+1)*B))
    rng_cfg = 2
    Xh = np.empty((B, N, I, J), dtype=np.complex64)
    for b in range(B):
        Xh[b] = make_mixture(N, I, J, seed=1000 * rng_cfg + rank * B + b, mode="mix")
    T0, V0 = make_nmf_init(N, I, J, K, seed=42)
    X_pinned = torch.from_numpy(Xh).pin_memory()
    Xd = X_pinned.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def make_sep(**kw):
        m = GaussILRMA(n_basis=K, spatial_algorithm=wl["spatial"], record_loss=False, **kw)
        if args.modular:
            m.fast_path = False
        if args.chunk:
            m.chunk_size = args.chunk
        if args.streams:
            m.n_streams = args.streams
        return m

    # ---- device-resident throughput ("value") -------------------------------------------------
    sep = make_sep(scale_restoration=False)
    sep(Xd, n_iter=0, basis=T0, activation=V0)  # binds the plan; state stays on the device
    sep.run_iterations(warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    # K steps = K x update_once over every mixture of the batch (ssspy/bss/base.py:68-77); the engine runs
    # them chunk-major (mixtures are independent), see ssspy_b200/bss/_engine.py
    sep.run_iterations(steps)
    ev1.record()
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - n0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        dist.barrier()
    if rank == 0:
        # keep the identical load running while the clock sampler collects (the timed region itself can
        # be shorter than one NVML sampling period); these extra steps are not timed
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            sep.run_iterations(steps)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 1.5 s of the same steps"
    if world > 1:
        dist.barrier()
    ms_per_step = ms_total / steps
    value = B * world * steps / (ms_total / 1e3)

    # ---- per-kernel times (CUDA events after every launch on the launching stream) -------------
    # The timed region above runs the batch as several chunk plans on concurrent streams (engine default), where
    # event-to-event deltas of interleaved launches mean nothing; the per-kernel breakdown is therefore taken from
    # the same steps run as ONE plan on one stream (every launch then covers the whole per-GPU batch, which is also
    # what the algorithmic-bytes figure of the dominant kernel refers to).
    prof_steps = 5
    sep_prof = make_sep(scale_restoration=False)
    sep_prof.chunk_size = B
    sep_prof(Xd, n_iter=0, basis=T0, activation=V0)
    sep_prof.run_iterations(2)
    torch.cuda.synchronize()
    _lib.call("ssb_profile_begin", torch.cuda.current_stream().cuda_stream)
    for _ in range(prof_steps):
        sep_prof.update_once()
    kernels = _lib.profile_end()
    del sep_prof
    total_prof = sum(k[2] for k in kernels) or 1.0
    dom = kernels[0] if kernels else ("none", 0, 0.0)
    abytes_step = algorithmic_bytes_per_mixture_iteration(N, I, J, K) * B
    peak, peak_src = hbm_peak()
    achieved = abytes_step / (ms_per_step / 1e3) / 1e9
    # dominant kernel on its own: its compulsory bytes (what it must read/write even in a perfectly fused
    # iteration: X once + the small T/V/W state) over its average launch duration, and the DRAM traffic
    # ncu measured for one launch of it (profiles/r1_ncu_traffic.json, same workload)
    alg_basis = 8 * N * I * J + 4 * (2 * N * I * K + N * K * J)
    alg_cov = 8 * N * I * J + 4 * (N * I * K + N * K * J) + 8 * N * N * N * I
    alg_act = 4 * N * I * J + 4 * (N * I * K + 2 * N * K * J)
    dom_alg = {"fused_basis": alg_basis, "coop_basis": alg_basis, "fused_phi_cov": alg_cov, "coop_phi_cov": alg_cov,
               "fused_activation": alg_act, "coop_activation": alg_act}.get(dom[0])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            tj = json.load(f)
        if (N, I, J, K, B) == (2, 1025, 512, 16, 64):
            traffic = tj["dram_bytes_per_launch"].get({"coop_basis": "kf_basis_coop", "coop_activation": "kf_activation_coop",
                                                       "fused_phi_cov": "kf_phi_cov"}.get(dom[0], ""))
    except Exception:
        traffic = None
    dom_ms = dom[2] / max(dom[1], 1)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "scope": "achieved/frac: algorithmic bytes of one whole update_once step over the per-GPU batch "
                         "(SURVEY.md 8(d): %.1f MB, %d launches) / step time; traffic: ncu DRAM bytes of one launch of the "
                         "dominant kernel" % (abytes_step / 1e6, launches // steps),
                "dominant_kernel": dom[0], "dominant_kernel_ms": dom_ms,
                "dominant_kernel_share": dom[2] / total_prof,
                "dominant_kernel_achieved": (dom_alg * B / (dom_ms / 1e3) / 1e9) if dom_alg else None,
                "dominant_kernel_frac": (dom_alg * B / (dom_ms / 1e3) / 1e9 / peak) if dom_alg else None,
                "kernels_ms_per_step": {k[0]: round(k[2] / prof_steps, 4) for k in kernels}}

    # ---- end to end through the public API with host buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        def e2e_once():
            m = make_sep(scale_restoration=True)
            t0 = time.perf_counter()
            Y = m(X_pinned, n_iter=steps, basis=T0, activation=V0)  # H2D, iterate, projection back, D2H
            torch.cuda.synchronize()
            return time.perf_counter() - t0, Y
        e2e_once()
        if world > 1:
            dist.barrier()
        dt, Y = e2e_once()
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": B * world * steps / dt, "unit": "mixture-iterations/s",
               "h2d_bytes_per_step": int((Xh.nbytes + 4 * (T0.size + V0.size)) / steps),
               "d2h_bytes_per_step": int(Y.numel() * 8 / steps),
               "note": "GaussILRMA.__call__(pinned host complex64 tensor, n_iter=%d): H2D of X, %d update_once, "
                       "projection back, separate, D2H of Y into pinned memory" % (steps, steps)}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        workers = os.cpu_count() or 1
        rate, cores, sample = cpu_arm(wl, 2, workers)
        cpu_baseline = {"value": rate, "unit": "mixture-iterations/s", "cores": cores, "kind": "port", "sample": sample}

    line = {"metric": "mixture_iterations_per_sec", "value": value, "unit": "mixture-iterations/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 state, fp64 N x N solves)", "data": "synthetic",
            "config": config, "batch_iterations_per_sec": 1e3 / ms_per_step, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
