#!/usr/bin/env python
"""
bench.py -- GS iterations/sec at 4096^2 complex64 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): Hologram, SLM 1152x1920 zero-padded to 4096x4096, method
"WGS-Kim", 50 iterations per optimize().  One STEP = one optimize() of 50 iterations from a reset state
(first row pass + 50 fused iterations + _populate_results).  N > 1: every rank runs its own
independent hologram (weak scaling, no collective inside the loop) and the run ends with ONE
all-gather of the final phases.

Target.  SURVEY.md 8d defines two targets for this config: 64 unit spots (parity) and a dense random
target "for throughput only".  The library skips far-field column tiles whose weights are all zero
(identical results), which makes throughput depend on the target, so the HEADLINE
(`value`, `e2e`, `roofline`) is measured on the DENSE target, where every tile is processed -- the
data-independent worst case -- and the 64-spot target is reported beside it under `sparse_target`.

Timing: CUDA events on the library's own stream (slmgs_timer_*), barrier + synchronise on both
sides, max over ranks.  The per-iteration working set (fld rows 38 MB + weights/target/phase_ff 192 MB)
exceeds the 126 MB L2, so no explicit flush is needed between iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GS iterations/sec @ 4096^2 c64"
SHAPE = (4096, 4096)
SLM_SHAPE = (1152, 1920)
METHOD = "WGS-Kim"
ITERS = 50
N_SPOTS = 64
WORKLOAD = "Hologram 1920x1152 SLM padded to 4096x4096, WGS-Kim, 50 iters (BASELINE configs[1])"


def make_inputs(seed, kind="dense"):
    """dense random target (throughput variant) or 64 unit spots at default_rng(1)
    positions (parity variant); seeded explicit phase."""
    rng = np.random.default_rng(1)
    pts = rng.integers(0, SHAPE[0], (2, N_SPOTS))
    if kind == "spots":
        target = np.zeros(SHAPE, dtype=np.float32)
        target[pts[1], pts[0]] = 1
    else:
        target = rng.random(SHAPE, dtype=np.float32)
    phase = np.random.default_rng(1000 + seed).uniform(-np.pi, np.pi, SLM_SHAPE).astype(np.float32)
    return target, phase


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~5 ms from a thread
    (nvidia-smi -lms as a fallback)."""

    def __init__(self, device):
        self.device = device
        self.samples = []  # (t, sm_mhz, reasons bitmask)
        self.smax = None
        self.stop_flag = False
        self.thread = None
        self.mode = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # torchrun sets CUDA_VISIBLE_DEVICES per rank only if asked to; LOCAL_RANK indexes the visible list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(vis.split(",")[self.device]) if vis and vis.split(",")[self.device].isdigit() else self.device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self, t0, t1):
        if self.mode != "nvml":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in inside))
        return {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(inside)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port of the NumPy path) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, iters = args
    from oracle import gs_oracle

    target, phase = make_inputs(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = gs_oracle.OracleHologram(target, phase=phase, slm_shape=SLM_SHAPE)
        t0 = time.perf_counter()
        h.optimize(method=METHOD, maxiter=iters, verbose=False)
        return time.perf_counter() - t0


def cpu_it_per_s(procs, iters):
    """`procs` independent holograms, one per process (NumPy's pocketfft and ufuncs are single-threaded)."""
    if procs == 1:
        dt = _cpu_worker((0, iters))
        return iters / dt, dt
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, iters) for i in range(procs)])
        dt = time.perf_counter() - t0
    return procs * iters / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    iters = 2  # bounded sample: 2 WGS-Kim iterations (+ the trailing _populate_results transform) per hologram
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_it_per_s(procs, 1)
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        v, _dt = cpu_it_per_s(procs, iters)
        vals.append(v)
        if time.perf_counter() - t_all > 150:
            break
    value = float(np.median(vals))
    sample = (f"{procs} independent holograms in {procs} processes (NumPy is single-threaded), {iters} {METHOD} "
              f"iterations each at 4096^2 incl. the trailing _populate_results transform; median of {len(vals)} step(s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * procs * iters / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "method": METHOD, "shape": list(SHAPE), "slm_shape": list(SLM_SHAPE)},
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# this repo
# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch  # plumbing only: pinned host memory, torch.distributed, NCCL all-gather

    from slmsuite_b200 import Hologram, _lib

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
    fp = C.POINTER(C.c_float)

    class Workload:
        """One target variant of the bench config: a device-resident step and an end-to-end step."""

        def __init__(self, kind):
            target, phase0 = make_inputs(rank, kind)
            self.holo = Hologram(target, phase=phase0, slm_shape=SLM_SHAPE, device=local_rank)
            self.ctx = self.holo._ctx
            self.holo._check(lib.slmgs_save_phase(self.ctx))
            # End-to-end path: every step uploads its inputs (normalised target + initial phase) from pinned host
            # memory through the C ABI, optimises, and downloads the resulting phase.  Two holograms are kept in
            # flight (ping-pong) so the copies of one step overlap the kernels of the other, the way a caller that
            # streams holograms through the public API would use it; each context has its own stream.
            self.holo_b = Hologram(target, phase=phase0, slm_shape=SLM_SHAPE, device=local_rank)
            self.pin_target = torch.from_numpy(np.ascontiguousarray(self.holo.target)).pin_memory()
            self.pin_phase = torch.from_numpy(phase0).pin_memory()
            self.pin_out = [torch.empty(SLM_SHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
            self.p_target = C.cast(self.pin_target.data_ptr(), fp)
            self.p_phase = C.cast(self.pin_phase.data_ptr(), fp)
            self.p_out = [C.cast(t.data_ptr(), fp) for t in self.pin_out]

        def step_resident(self):
            """inputs already in HBM: restore phase + weights on the device, then one optimize()."""
            holo = self.holo
            holo._check(lib.slmgs_restore_phase(self.ctx))
            holo.reset(reset_phase=False)
            holo.flags["fixed_phase"] = False
            holo.optimize(METHOD, maxiter=ITERS, verbose=False)

        def e2e_submit(self, h):
            """upload target + phase (host buffers), reset, launch optimize() asynchronously."""
            h._check(lib.slmgs_set_target(h._ctx, self.p_target, 0))
            h._check(lib.slmgs_set_phase(h._ctx, self.p_phase))
            h.reset(reset_phase=False)
            h.flags["fixed_phase"] = False
            h.optimize(METHOD, maxiter=ITERS, verbose=False)

        def e2e_collect(self, h, slot):
            h._check(lib.slmgs_get_phase(h._ctx, self.p_out[slot]))  # D2H, synchronises that hologram's stream

        def run_e2e(self, n_steps):
            pair = (self.holo, self.holo_b)
            self.e2e_submit(pair[0])
            for i in range(n_steps):
                if i + 1 < n_steps:
                    self.e2e_submit(pair[(i + 1) % 2])
                self.e2e_collect(pair[i % 2], i % 2)

        def sync(self):
            self.holo._check(lib.slmgs_sync(self.ctx))
            self.holo_b._check(lib.slmgs_sync(self.holo_b._ctx))

    wl = Workload("dense")
    holo, ctx, chk = wl.holo, wl.ctx, wl.holo._check
    step_resident, run_e2e = wl.step_resident, wl.run_e2e

    def barrier():
        torch.cuda.synchronize()  # device-wide: covers the library's own streams
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # the final all-gather of phases (one per job), over the library's device buffer
    class _DevPhase:
        def __init__(self):
            self.__cuda_array_interface__ = {
                "shape": SLM_SHAPE, "typestr": "<f4", "data": (lib.slmgs_phase_device_ptr(ctx), False), "version": 3}

    def allgather_phases():
        if dist is None:
            return 0.0
        src = torch.as_tensor(_DevPhase(), device=torch.device("cuda", local_rank))
        out = torch.empty((world,) + SLM_SHAPE, dtype=torch.float32, device=src.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(out, src)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_resident()
    run_e2e(2)
    allgather_phases()
    barrier()

    # ---- timed region: K steps, device events on the launching stream ---------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)
    launches0 = lib.slmgs_launch_count(ctx)
    barrier()
    w0 = time.perf_counter()
    chk(lib.slmgs_timer_start(ctx))
    host_s = 0.0
    for _ in range(args.steps):
        h0 = time.perf_counter()
        step_resident()
        host_s += time.perf_counter() - h0
    ms = C.c_float()
    chk(lib.slmgs_timer_stop(ctx, C.byref(ms)))
    ag_ms = allgather_phases()
    barrier()
    w1 = time.perf_counter()
    launches = lib.slmgs_launch_count(ctx) - launches0
    clocks = sampler.stop(w0, w1)
    total_ms = max_over_ranks(float(ms.value) + ag_ms)
    wall_ms = max_over_ranks(1e3 * (w1 - w0))
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- per-kernel durations: the same K steps again with CUDA events around every launch ----------
    # (event records between kernels switch off programmatic dependent launch, so they stay out of the timed
    # region above; same process, same inputs, immediately afterwards)
    chk(lib.slmgs_profile_enable(ctx, 1))
    for _ in range(args.steps):
        step_resident()
    prof_ms = (C.c_float * 6)()
    prof_n = (C.c_int * 6)()
    chk(lib.slmgs_profile_read(ctx, prof_ms, prof_n))
    chk(lib.slmgs_profile_enable(ctx, 0))

    # ---- end-to-end: same steps through host buffers ---------------------------------------------
    barrier()
    e0 = time.perf_counter()
    run_e2e(args.steps)
    wl.sync()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - e0))  # host clock: copies are synchronous C-ABI calls
    barrier()
    dense_info = holo.sparse_info()

    # ---- the 64-spot target of the same config (sparse far field: most column tiles are skipped) --------
    ws = Workload("spots")
    for _ in range(3):
        ws.step_resident()
    ws.run_e2e(2)
    ws.sync()
    barrier()
    sp_launch0 = lib.slmgs_launch_count(ws.ctx)
    ws.holo._check(lib.slmgs_timer_start(ws.ctx))
    for _ in range(args.steps):
        ws.step_resident()
    sp_ms = C.c_float()
    ws.holo._check(lib.slmgs_timer_stop(ws.ctx, C.byref(sp_ms)))
    sp_launches = lib.slmgs_launch_count(ws.ctx) - sp_launch0
    sp_total_ms = max_over_ranks(float(sp_ms.value))
    barrier()
    e0 = time.perf_counter()
    ws.run_e2e(args.steps)
    ws.sync()
    sp_e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - e0))
    barrier()
    sp_info = ws.holo.sparse_info()
    ws.holo._check(lib.slmgs_profile_enable(ws.ctx, 1))
    for _ in range(args.steps):
        ws.step_resident()
    sp_prof_ms = (C.c_float * 6)()
    sp_prof_n = (C.c_int * 6)()
    ws.holo._check(lib.slmgs_profile_read(ws.ctx, sp_prof_ms, sp_prof_n))
    ws.holo._check(lib.slmgs_profile_enable(ws.ctx, 0))

    # ---- north_star's own yardstick: the fused GS iteration on a dense 4096^2 field (slm_shape == shape) -------
    gs_dense = None
    if rank == 0:
        rng = np.random.default_rng(7)
        hd = Hologram(rng.random(SHAPE, dtype=np.float32), phase=rng.uniform(-np.pi, np.pi, SHAPE).astype(np.float32),
                      device=local_rank)
        hd.optimize("GS", maxiter=ITERS, verbose=False)
        hd._check(lib.slmgs_sync(hd._ctx))
        hd._check(lib.slmgs_timer_start(hd._ctx))
        reps = 3
        for _ in range(reps):
            hd.optimize("GS", maxiter=ITERS, verbose=False)
        gs_ms = C.c_float()
        hd._check(lib.slmgs_timer_stop(hd._ctx, C.byref(gs_ms)))
        gs_its = reps * ITERS / (gs_ms.value * 1e-3)
        gs_dense = {"it_per_s": gs_its, "ms_per_iteration": gs_ms.value / (reps * ITERS),
                    "what": "Hologram 4096x4096 with slm_shape == shape (no zero padding), dense random target, method GS, "
                            f"{reps} x optimize(maxiter={ITERS}) incl. the trailing _populate_results transform, CUDA events"}
        del hd
    barrier()

    iters_total = world * args.steps * ITERS
    value = iters_total / (total_ms * 1e-3)
    e2e_value = iters_total / (e2e_ms * 1e-3)
    sparse_target = {
        "target": f"{N_SPOTS} unit spots (SURVEY.md 8d config 2, parity variant)",
        "value": iters_total / (sp_total_ms * 1e-3), "unit": "it/s", "ms_per_step": sp_total_ms / args.steps,
        "e2e": {"value": iters_total / (sp_e2e_ms * 1e-3), "unit": "it/s", "ms_per_step": sp_e2e_ms / args.steps},
        "gpu_launches": int(sp_launches),
        "kernels": {n: {"launches": int(sp_prof_n[k]), "avg_ms": float(sp_prof_ms[k]) / int(sp_prof_n[k])}
                    for k, n in enumerate(["row_first", "row_fused", "row_last", "col_forward", "col_fused", "col_inverse"])
                    if sp_prof_n[k]},
        "sparse_path_used": bool(sp_info[0]), "active_column_tiles": int(sp_info[1]), "column_tiles": int(sp_info[2]),
        "note": "same config and code path; column tiles whose weights are all zero are skipped (identical results, "
                "tests/test_sparse.py); not the headline because it depends on the target",
    }

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (column fused) -------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    P = SHAPE[0] * SHAPE[1]
    hW = SLM_SHAPE[0] * SHAPE[1]
    # model bytes per launch (DESIGN.md "Algorithmic bytes"): column kernel = forward + inverse column pass
    # (4 x 8P) + weights 4P + target 4P + weights write 4P (WGS) [+ phase_ff 4P once Kim has fixed the phase];
    # averaged over the 50 iterations of this workload: iteration 0 has no update (36 P), iterations 1-8 update with
    # the phase taken from the field (44 P), iterations 9-49 use the stored phase (48 P; iteration 9 stores it with one
    # extra forward column pass, counted under col_forward).
    col_model = (1 * 36 + 8 * 44 + 41 * 48) / 50.0 * P
    row_model = 32.0 * P
    # bytes the implementation must actually move (zero-padding skipped: only the h SLM rows of fld are touched)
    col_actual = 16.0 * hW + (1 * 4 + 8 * 12 + 41 * 16) / 50.0 * P
    row_actual = 16.0 * hW
    kern = {}
    names = ["row_first", "row_fused", "row_last", "col_forward", "col_fused", "col_inverse"]
    for k in range(6):
        if prof_n[k]:
            kern[names[k]] = {"launches": int(prof_n[k]), "avg_ms": float(prof_ms[k]) / int(prof_n[k])}
    share = {k: v["avg_ms"] * v["launches"] for k, v in kern.items()}
    tot = sum(share.values()) or 1.0
    dom = "col_fused" if share.get("col_fused", 0) >= share.get("row_fused", 0) else "row_fused"
    dom_ms = kern[dom]["avg_ms"]
    model = col_model if dom == "col_fused" else row_model
    actual = col_actual if dom == "col_fused" else row_actual
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": model / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": model / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
        "model_bytes_per_launch": model, "avg_launch_ms": dom_ms,
        "actual_bytes_per_launch": actual, "actual_gbs": actual / (dom_ms * 1e-3) / 1e9,
        "actual_frac": actual / (dom_ms * 1e-3) / 1e9 / peak,
        "kernel_share_of_step": {k: v / tot for k, v in share.items()},
        "measured": "CUDA events around every launch of the same K steps, repeated right after the timed region "
                    "(event records between kernels disable programmatic dependent launch)",
        "kernels": kern,
        "iteration_model_frac": (68.0 + 76.0 * 8 + 80.0 * 41) / 50.0 * P * (value / world) / 1e9 / peak,
    }
    if gs_dense is not None:
        # 68 P bytes per fused GS iteration against the measured copy bandwidth: north_star's ">= 60 %"
        gs_dense["model_gbs"] = 68.0 * P * gs_dense["it_per_s"] / 1e9
        gs_dense["model_frac_of_hbm_peak"] = gs_dense["model_gbs"] / peak

    # ---- CPU baseline (oracle port of the reference's NumPy path), bounded sample ------------------
    cpu_iters = 2
    cpu_value, cpu_dt = cpu_it_per_s(1, cpu_iters)
    cpu = {"value": cpu_value, "unit": "it/s", "cores": 1, "kind": "port",
           "sample": f"{cpu_iters} {METHOD} iterations at 4096^2 (+ trailing _populate_results transform), oracle port, "
                     f"1 process ({os.cpu_count()} host cores present), {cpu_dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "method": METHOD, "iters_per_step": ITERS, "shape": list(SHAPE),
                   "slm_shape": list(SLM_SHAPE), "parallelism": f"replicas x{world}",
                   "target": "dense random (SURVEY.md 8d config 2 throughput variant): every far-field column tile "
                             f"is processed ({dense_info[1]}/{dense_info[2]} active, sparse path used: {bool(dense_info[0])})",
                   "l2": "working set 230 MB/iteration > 126 MB L2, no flush needed",
                   "final_allgather_ms": ag_ms},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": int(4 * P + 4 * SLM_SHAPE[0] * SLM_SHAPE[1]),
                "d2h_bytes_per_step": int(4 * SLM_SHAPE[0] * SLM_SHAPE[1]), "ms_per_step": e2e_ms / args.steps,
                "how": "C-ABI calls with pinned host buffers; two holograms in flight so copies overlap kernels"},
        "gpu_launches": total_launches,
        "wall_ms_per_step": wall_ms / args.steps,
        "host_submit_ms_per_step": 1e3 * host_s / args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sparse_target": sparse_target,
        "gs_dense_4096": gs_dense,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_STDOUT_FD = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    # rank 0 prints ONE JSON line on stdout: everything else that writes to fd 1 during the run (NCCL's version
    # banner, library chatter) is sent to stderr by pointing fd 1 at fd 2 until the line is emitted
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
