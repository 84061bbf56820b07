#!/usr/bin/env python
"""
bench.py -- GS iterations/sec at 4096^2 complex64 (BASELINE.json metric), HBM GB/s against the roofline.

    python bench.py --gpus N --steps K --warmup W            # this repo (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the unmodified reference's NumPy path on the host cores

HEADLINE workload = the configuration the metric is quoted on (BASELINE.md "GS 4096^2, slm 4096^2 (headline shape)",
the loop slmsuite/holography/algorithms/_hologram.py:1465-1490): ``Hologram`` with ``shape == slm_shape == 4096^2``
(no zero padding), dense random target, method "GS", 50 iterations per ``optimize()``.  One STEP = one
``optimize("GS", maxiter=50)`` from a restored initial phase (first row pass + 50 fused iterations +
``_populate_results``).  ``value`` = iterations/s with everything resident in HBM (CUDA events on the library's
stream); ``e2e`` = the same through the C ABI with pinned HOST buffers (target + phase up, phase down, every step).
N > 1: every rank runs its own hologram (weak scaling, no collective inside the loop); the run ends with ONE
all-gather of the final phases.

``roofline`` (dominant kernel = fused column kernel): bytes the launch MUST move -- field read + write 16 P plus the
weights image 4 P = 20 P (row kernel: 16 P) -- over its CUDA-event duration, against MEASURED_PEAKS.json ``hbm_gbs``.
``model_frac_68P`` is SURVEY.md 8d's four-pass contract model (68 P bytes per GS iteration) for the whole iteration.

The other BASELINE configs are reported in named extra keys: ``config2_padded_kim`` (1152x1920 in 4096^2, WGS-Kim;
dense and 64-spot target), ``config3_spot_feedback``, ``config5_8192`` (a bounded sample of its 200 iterations),
``config1_512`` and the reference's own benchmark shape ``refbench_1024``; N > 1 adds ``config4_sharded`` (64 x 2048^2
GS sharded over the ranks with the final all-gather timed separately).

Timing: CUDA events on the library's own stream (slmgs_timer_*), barrier + synchronise on both sides, max over
ranks.  The headline's working set per iteration (field 134 MB + weights 67 MB) exceeds the 126 MB L2, so no explicit
flush is needed between iterations.
"""
import argparse
import ctypes as C
import json
import os
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
METHOD = "GS"
ITERS = 50
WORKLOAD = ("Hologram 4096x4096, slm_shape == shape (dense field, no zero padding), dense random target, method GS, "
            "50 iterations per optimize() (BASELINE metric configuration)")
PADDED_SLM = (1152, 1920)
N_SPOTS = 64


def make_inputs(seed, shape=SHAPE, slm=None, kind="dense", n_spots=N_SPOTS):
    """Seeded synthetic inputs: dense random target or unit spots; explicit phase (never an RNG of the
    implementation under test)."""
    slm = tuple(slm or shape)
    rng = np.random.default_rng(1)
    pts = rng.integers(0, shape[0], (2, n_spots))
    if kind == "spots":
        target = np.zeros(shape, dtype=np.float32)
        target[pts[1], pts[0]] = 1
    else:
        target = rng.random(shape, dtype=np.float32)
    phase = np.random.default_rng(1000 + seed).uniform(-np.pi, np.pi, slm).astype(np.float32)
    return target, phase


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every ~4 ms from a thread."""

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
# reference arm: the reference's own CPU implementation (NumPy backend) on the host cores
# ------------------------------------------------------------------------------------------------
def _reference_class():
    """(Hologram class, kind): the unmodified reference from baseline/_ref, else the oracle port."""
    from baseline import ref_arm

    if ref_arm.available():
        return ref_arm.load().Hologram, "reference"
    from oracle import gs_oracle  # the one other place bench.py may execute oracle/: the reference arm

    return gs_oracle.OracleHologram, "port"


def _cpu_worker(args):
    """One hologram of the headline workload on one core; returns (seconds, kind[, amp_ff])."""
    seed, iters, want_result = args
    cls, kind = _reference_class()
    target, phase = make_inputs(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = cls(target, phase=phase.copy())
        t0 = time.perf_counter()
        h.optimize(method=METHOD, maxiter=iters, verbose=False)
        dt = time.perf_counter() - t0
    return (dt, kind, np.array(h.amp_ff, dtype=np.float32)) if want_result else (dt, kind, None)


def cpu_it_per_s(procs, iters):
    """`procs` independent holograms, one per process (NumPy's pocketfft and ufuncs are single-threaded)."""
    if procs == 1:
        dt, kind, res = _cpu_worker((0, iters, True))
        return iters / dt, dt, kind, res
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        out = pool.map(_cpu_worker, [(i, iters, False) for i in range(procs)])
        dt = time.perf_counter() - t0
    return procs * iters / dt, dt, out[0][1], None


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    iters = 2  # bounded sample: 2 GS iterations (+ the trailing _populate_results transform) per hologram
    vals = []
    kind = "reference"
    if args.warmup >= 1:
        cpu_it_per_s(procs, 1)
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        v, _dt, kind, _ = cpu_it_per_s(procs, iters)
        vals.append(v)
        if time.perf_counter() - t_all > 150:
            break
    value = float(np.median(vals))
    sample = (f"{procs} independent holograms in {procs} processes (NumPy is single-threaded), {iters} {METHOD} "
              f"iterations each at 4096^2 incl. the trailing _populate_results transform; median of {len(vals)} step(s); "
              + ("unmodified reference from baseline/_ref, NumPy backend" if kind == "reference" else "oracle port"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * procs * iters / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "method": METHOD, "iters_per_step": ITERS, "shape": list(SHAPE),
                   "slm_shape": list(SHAPE)},
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# this repo
# ------------------------------------------------------------------------------------------------
KERNEL_NAMES = ["row_first", "row_fused", "row_last", "col_forward", "col_fused", "col_inverse"]


def run_b200(args, rank, local_rank, world):
    import torch  # plumbing only: pinned host memory and a device-wide synchronise

    from slmsuite_b200 import Hologram, _lib
    from slmsuite_b200 import comm as slm_comm

    lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
    fp = C.POINTER(C.c_float)
    # one process per GPU: host threads and pinned buffers on the cores next to this rank's GPU (matters for the
    # end-to-end copies on a two-socket box; SLMGS_BENCH_NUMA=0 switches it off for A/B runs)
    numa_cpus = _lib.bind_host_to_device(local_rank) if os.environ.get("SLMGS_BENCH_NUMA", "1") != "0" else None
    torch.cuda.set_device(local_rank)
    # one process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*): the package's own communicator --
    # TCP rendezvous + NCCL loaded inside libslmgs.so (slmgs_comm_* / slmgs_allgather_phase); no torch.distributed
    cm = slm_comm.Comm(rank, world, os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")),
                       device=local_rank) if world > 1 else None
    if cm is not None:
        slm_comm.set_default(cm)

    def barrier():
        torch.cuda.synchronize()  # device-wide: covers the library's own streams
        if cm is not None:
            cm.barrier()

    def max_over_ranks(x):
        if cm is None:
            return x
        return float(max(float(v[0]) for v in cm.allgather_host(np.array([x], dtype=np.float64))))

    def sum_over_ranks(x):
        if cm is None:
            return x
        return float(sum(float(v[0]) for v in cm.allgather_host(np.array([x], dtype=np.float64))))

    class Workload:
        """One hologram configuration: a device-resident step and an end-to-end step."""

        def __init__(self, method, iters, shape=SHAPE, slm=None, kind="dense", e2e=True, **kw):
            self.method, self.iters, self.kw = method, iters, kw
            self.slm = tuple(slm or shape)
            target, phase0 = make_inputs(rank, shape, self.slm, kind)
            self.holo = Hologram(target, phase=phase0, slm_shape=self.slm, device=local_rank)
            self.ctx = self.holo._ctx
            self.holo._check(lib.slmgs_save_phase(self.ctx))
            self.P = shape[0] * shape[1]
            self.S = self.slm[0] * self.slm[1]
            if e2e:
                # End-to-end path: every step uploads its inputs (normalised target + initial phase) from pinned host
                # memory through the C ABI, optimises, and downloads the resulting phase.  Two holograms are kept in
                # flight (ping-pong) so the copies of one step overlap the kernels of the other, the way a caller that
                # streams holograms through the public API would use it; each context has its own stream.
                self.holo_b = Hologram(target, phase=phase0, slm_shape=self.slm, device=local_rank)
                self.pin_target = torch.from_numpy(np.ascontiguousarray(self.holo.target)).pin_memory()
                self.pin_phase = torch.from_numpy(phase0).pin_memory()
                self.pin_out = [torch.empty(self.slm, dtype=torch.float32).pin_memory() for _ in range(2)]
                self.p_target = C.cast(self.pin_target.data_ptr(), fp)
                self.p_phase = C.cast(self.pin_phase.data_ptr(), fp)
                self.p_out = [C.cast(t.data_ptr(), fp) for t in self.pin_out]

        def step_resident(self):
            """inputs already in HBM: restore phase + weights on the device, then one optimize()."""
            holo = self.holo
            holo._check(lib.slmgs_restore_phase(self.ctx))
            holo.reset(reset_phase=False)
            holo.flags["fixed_phase"] = False
            holo.optimize(self.method, maxiter=self.iters, verbose=False, **self.kw)

        def e2e_submit(self, h):
            """upload target + phase (host buffers), reset, launch optimize() asynchronously."""
            h._check(lib.slmgs_set_target(h._ctx, self.p_target, 0))
            h._check(lib.slmgs_set_phase(h._ctx, self.p_phase))
            h.reset(reset_phase=False)
            h.flags["fixed_phase"] = False
            h.optimize(self.method, maxiter=self.iters, verbose=False, **self.kw)

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
            if hasattr(self, "holo_b"):
                self.holo_b._check(lib.slmgs_sync(self.holo_b._ctx))

        def timed(self, steps):
            """K device-timed steps (CUDA events on the library's stream); returns (ms, launches, host seconds)."""
            n0 = lib.slmgs_launch_count(self.ctx)
            self.holo._check(lib.slmgs_timer_start(self.ctx))
            host = 0.0
            for _ in range(steps):
                h0 = time.perf_counter()
                self.step_resident()
                host += time.perf_counter() - h0
            ms = C.c_float()
            self.holo._check(lib.slmgs_timer_stop(self.ctx, C.byref(ms)))
            return float(ms.value), lib.slmgs_launch_count(self.ctx) - n0, host

        def profile(self, steps):
            """per-kernel durations: the same steps again with CUDA events around every launch (event records between
            kernels switch off programmatic dependent launch, so they stay out of the timed region)."""
            self.holo._check(lib.slmgs_profile_enable(self.ctx, 1))
            for _ in range(steps):
                self.step_resident()
            pms = (C.c_float * 6)()
            pn = (C.c_int * 6)()
            self.holo._check(lib.slmgs_profile_read(self.ctx, pms, pn))
            self.holo._check(lib.slmgs_profile_enable(self.ctx, 0))
            return {n: {"launches": int(pn[k]), "avg_ms": float(pms[k]) / int(pn[k])} for k, n in enumerate(KERNEL_NAMES)
                    if pn[k]}

        def close(self):
            self.sync()
            for name in ("holo", "holo_b"):
                if hasattr(self, name):
                    delattr(self, name)

    def side_config(label, wl, steps, e2e_steps=0, extra=None):
        """A named extra key: device-timed value (+ e2e) of another BASELINE config, max over ranks."""
        for _ in range(2):
            wl.step_resident()
        wl.sync()
        barrier()
        ms, launches, _ = wl.timed(steps)
        ms = max_over_ranks(ms)
        out = {"what": label, "value": world * steps * wl.iters / (ms * 1e-3), "unit": "it/s",
               "ms_per_step": ms / steps, "steps": steps, "iters_per_step": wl.iters, "gpu_launches": int(launches)}
        if e2e_steps:
            wl.run_e2e(2)
            wl.sync()
            barrier()
            t0 = time.perf_counter()
            wl.run_e2e(e2e_steps)
            wl.sync()
            e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
            out["e2e"] = {"value": world * e2e_steps * wl.iters / (e_ms * 1e-3), "unit": "it/s",
                          "ms_per_step": e_ms / e2e_steps}
        used, n_active, n_tiles = wl.holo.sparse_info()
        out["sparse_far_field"] = {"used": bool(used), "active_column_tiles": int(n_active), "column_tiles": int(n_tiles)}
        out["kernels"] = wl.profile(max(1, steps // 2))
        if extra:
            out.update(extra)
        return out

    # ================================ headline: dense GS 4096^2 ===================================
    wl = Workload(METHOD, ITERS)
    holo, ctx, chk = wl.holo, wl.ctx, wl.holo._check
    P = wl.P

    # the final all-gather of phases (one per job): NCCL behind the C ABI, on the hologram's own stream
    def allgather_phases():
        if cm is None:
            return 0.0
        cm.allgather_phase(holo, per_rank=1, download=False)
        return float(cm.last_allgather_ms)

    # ---- parity sample (rank 0): 2 iterations from the seeded inputs, compared below with the CPU reference ----
    parity_amp = None
    if rank == 0:
        wl.holo._check(lib.slmgs_restore_phase(ctx))
        holo.reset(reset_phase=False)
        holo.optimize(METHOD, maxiter=2, verbose=False)
        parity_amp = np.array(holo.amp_ff, dtype=np.float32)

    # ---- warm-up ------------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        wl.step_resident()
    wl.run_e2e(2)
    allgather_phases()
    barrier()

    # ---- timed region: K steps, device events on the launching stream ---------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)
    barrier()
    w0 = time.perf_counter()
    ms, launches, host_s = wl.timed(args.steps)
    ag_ms = allgather_phases()
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1)
    total_ms = max_over_ranks(ms + ag_ms)
    wall_ms = max_over_ranks(1e3 * (w1 - w0))
    total_launches = int(sum_over_ranks(float(launches)))

    kern = wl.profile(args.steps)
    # host cost of submitting ONE step into an empty queue (inside the timed region the host runs ahead until the launch
    # queue is full and then blocks in the launches, so host seconds per step there tend to the device time)
    wl.sync()
    h0 = time.perf_counter()
    wl.step_resident()
    host_idle_queue_ms = 1e3 * (time.perf_counter() - h0)
    wl.sync()

    # ---- end-to-end: same steps through host buffers ---------------------------------------------
    barrier()
    e0 = time.perf_counter()
    wl.run_e2e(args.steps)
    wl.sync()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - e0))  # host clock: copies are synchronous C-ABI calls
    barrier()
    head_info = holo.sparse_info()
    geom = (C.c_int * 4)()
    try:
        lib.slmgs_launch_geometry(ctx, geom)
        geometry = {"row_threads": int(geom[0]), "row_blocks": int(geom[1]), "col_threads": int(geom[2]),
                    "col_blocks": int(geom[3])}
    except Exception:
        geometry = None
    wl.close()
    del holo

    iters_total = world * args.steps * ITERS
    value = iters_total / (total_ms * 1e-3)
    e2e_value = iters_total / (e2e_ms * 1e-3)

    # ================================ the other BASELINE configs ===================================
    extras = {}
    side_steps = max(2, min(args.steps, 5))

    def try_extra(key, fn):
        try:
            extras[key] = fn()
        except Exception as exc:  # an extra key never takes the headline down
            extras[key] = {"error": f"{type(exc).__name__}: {exc}"}
        barrier()

    def cfg2():
        out = {}
        for kind, label in (("dense", "dense random target (every column tile processed)"),
                            ("spots", f"{N_SPOTS} unit spots (SURVEY.md 8d config 2; sparse far field)")):
            w2 = Workload("WGS-Kim", 50, SHAPE, PADDED_SLM, kind)
            out["dense_target" if kind == "dense" else "spot_target"] = side_config(label, w2, side_steps, e2e_steps=side_steps)
            w2.close()
        out["what"] = "BASELINE configs[1]: Hologram 1920x1152 SLM padded to 4096x4096, WGS-Kim, 50 iterations"
        return out

    def cfg3():
        from slmsuite_b200 import SpotHologram

        h = SpotHologram.make_rectangular_array(SHAPE, array_shape=(32, 32), array_pitch=(64, 64), basis="knm",
                                                device=local_rank)
        ph = np.random.default_rng(3000 + rank).uniform(-np.pi, np.pi, SHAPE).astype(np.float32)
        kw = dict(method="WGS-Leonardo", maxiter=100, feedback="computational_spot", verbose=False)
        h.reset_phase(ph)
        h._check(lib.slmgs_save_phase(h._ctx))
        h.optimize(**kw)
        h._check(lib.slmgs_sync(h._ctx))
        barrier()
        reps = 2
        h._check(lib.slmgs_timer_start(h._ctx))
        for _ in range(reps):  # a device-resident step, as the headline's: restore the initial phase, reset the weights, optimize
            h._check(lib.slmgs_restore_phase(h._ctx))
            h.reset(reset_phase=False)
            h.optimize(**kw)
        t = C.c_float()
        h._check(lib.slmgs_timer_stop(h._ctx, C.byref(t)))
        t_ms = max_over_ranks(float(t.value))
        used, n_active, n_tiles = h.sparse_info()
        return {"what": "BASELINE configs[2]: SpotHologram 32x32 spots on 4096x4096, WGS-Leonardo, computational_spot "
                        "feedback, 100 iterations per step (device-resident step: phase restore + weights reset + optimize)",
                "value": world * reps * 100 / (t_ms * 1e-3), "unit": "it/s", "ms_per_step": t_ms / reps, "steps": reps,
                "sparse_far_field": {"used": bool(used), "active_column_tiles": int(n_active), "column_tiles": int(n_tiles)}}

    def cfg5():
        from slmsuite_b200 import SpotHologram

        v = np.random.default_rng(5).uniform(64, 8192 - 64, (2, 10000))
        t0 = time.perf_counter()
        h = SpotHologram((8192, 8192), v, basis="knm", device=local_rank)
        ctor_s = time.perf_counter() - t0
        h.reset_phase(np.random.default_rng(5000 + rank).uniform(-np.pi, np.pi, (8192, 8192)).astype(np.float32))
        kw = dict(method="WGS-Leonardo", maxiter=20, feedback="computational_spot", verbose=False)
        h.optimize(**kw)
        h._check(lib.slmgs_sync(h._ctx))
        barrier()
        reps = 2
        h._check(lib.slmgs_timer_start(h._ctx))
        for _ in range(reps):
            h.optimize(**kw)
        t = C.c_float()
        h._check(lib.slmgs_timer_stop(h._ctx, C.byref(t)))
        t_ms = max_over_ranks(float(t.value))
        used, n_active, n_tiles = h.sparse_info()
        return {"what": "BASELINE configs[4]: SpotHologram 10k random spots, 8192x8192, WGS-Leonardo, computational_spot "
                        f"feedback; bounded sample: {reps} x optimize(maxiter=20) of the 200 iterations, one replica per GPU",
                "value": world * reps * 20 / (t_ms * 1e-3), "unit": "it/s", "ms_per_step": t_ms / reps, "steps": reps,
                "constructor_s": ctor_s,
                "sparse_far_field": {"used": bool(used), "active_column_tiles": int(n_active), "column_tiles": int(n_tiles)}}

    def cfg1():
        w1_ = Workload("GS", 30, (512, 512), None, "dense", e2e=False)
        out = side_config("BASELINE configs[0]: Hologram 512x512, dense random target, GS, 30 iterations", w1_, 20)
        w1_.close()
        return out

    def refbench():
        out = {"what": "the reference's own benchmark shape (tests/holography/test_algorithms.py:121-145): 1024x1024, "
                       "20 spots, 20 iterations"}
        for method in ("GS", "WGS-Leonardo", "WGS-Kim"):
            wr = Workload(method, 20, (1024, 1024), None, "spots", e2e=False)
            out[method] = side_config(method, wr, 10)
            wr.close()
        return out

    def cfg4():
        from slmsuite_b200 import HologramBatch

        n_total = 64
        per = -(-n_total // world)
        lo, hi = min(rank * per, n_total), min(rank * per + per, n_total)
        shp = (2048, 2048)
        T = np.zeros((hi - lo,) + shp, dtype=np.float32)
        Ph = np.empty((hi - lo,) + shp, dtype=np.float32)
        for i, b in enumerate(range(lo, hi)):
            r = np.random.default_rng(100 + b)
            T[i][r.integers(0, shp[0], 100), r.integers(0, shp[1], 100)] = 1
            Ph[i] = np.random.default_rng(200 + b).uniform(-np.pi, np.pi, shp)
        hb = HologramBatch(T, phase=Ph, device=local_rank)
        hb._check(lib.slmgs_save_phase(hb._ctx))
        hb.optimize("GS", maxiter=50, verbose=False)
        hb._check(lib.slmgs_sync(hb._ctx))
        barrier()
        reps = 2
        hb._check(lib.slmgs_timer_start(hb._ctx))
        for _ in range(reps):
            hb._check(lib.slmgs_restore_phase(hb._ctx))
            hb.reset(reset_phase=False)
            hb.optimize("GS", maxiter=50, verbose=False)
        t = C.c_float()
        hb._check(lib.slmgs_timer_stop(hb._ctx, C.byref(t)))
        loop_ms = max_over_ranks(float(t.value)) / reps
        barrier()
        t0 = time.perf_counter()
        phases = hb.gather_phases(n_total=n_total)  # the ONE collective of the job: 64 x 2048^2 f32 = 1 GiB on every rank
        gather_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        gather_dev_ms = max_over_ranks(float(cm.last_allgather_ms)) if cm is not None else 0.0
        used, n_active, n_tiles = hb.sparse_info()
        return {"what": "BASELINE configs[3]: 64 independent 2048x2048 Holograms (100 unit spots each), GS 50 iterations, "
                        f"sharded {hi - lo} per GPU, one all-gather of the final phases",
                "value": n_total * 50 / (loop_ms * 1e-3), "unit": "hologram-it/s (aggregate)", "loop_ms": loop_ms,
                "per_gpu_value": n_total * 50 / (loop_ms * 1e-3) / world,
                "allgather_ms": gather_ms, "allgather_device_ms": gather_dev_ms,
                "allgather_bytes": int(n_total * shp[0] * shp[1] * 4),
                "allgather_includes": "allgather_ms: NCCL all-gather (slmgs_allgather_phase) + D2H of the 1 GiB result into "
                                      "pageable host memory; allgather_device_ms: the collective alone (CUDA events)",
                "gathered_shape": list(np.shape(phases)),
                "sparse_far_field": {"used": bool(used), "active_column_tiles": int(n_active), "column_tiles": int(n_tiles)}}

    if args.extras:
        try_extra("config2_padded_kim", cfg2)
        try_extra("config3_spot_feedback", cfg3)
        try_extra("config4_sharded", cfg4)
        try_extra("config5_8192", cfg5)
        if world == 1:
            try_extra("config1_512", cfg1)
            try_extra("refbench_1024", refbench)

    if rank != 0:
        if cm is not None:
            cm.barrier()
            cm.close()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # bytes a launch must move: the two-pass design reads and writes the c64 field once per kernel
    bytes_must = {"col_fused": 20.0 * P, "row_fused": 16.0 * P}
    share = {k: v["avg_ms"] * v["launches"] for k, v in kern.items()}
    tot = sum(share.values()) or 1.0
    dom = "col_fused" if share.get("col_fused", 0) >= share.get("row_fused", 0) else "row_fused"
    dom_ms = kern[dom]["avg_ms"]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dense_gs_4096", {}).get(dom)
    except Exception:
        pass
    per_kernel = {k: {"bytes_per_launch": bytes_must[k], "avg_launch_ms": kern[k]["avg_ms"],
                      "gbs": bytes_must[k] / (kern[k]["avg_ms"] * 1e-3) / 1e9,
                      "frac": bytes_must[k] / (kern[k]["avg_ms"] * 1e-3) / 1e9 / peak}
                  for k in bytes_must if k in kern}
    it_ms = sum(kern[k]["avg_ms"] for k in bytes_must if k in kern)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": bytes_must[dom] / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": bytes_must[dom] / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
        "bytes_per_launch": bytes_must[dom], "avg_launch_ms": dom_ms,
        "bytes_model": "bytes the launch must move: c64 field read + write (16 P) + weights image read (4 P) = 20 P for the "
                       "fused column kernel, 16 P for the fused row kernel; P = 4096^2",
        "per_kernel": per_kernel,
        "iteration_frac_36P": 36.0 * P / (it_ms * 1e-3) / 1e9 / peak if it_ms else None,
        "model_frac_68P": 68.0 * P * (value / world) / 1e9 / peak,
        "model_frac_68P_note": "SURVEY.md 8d contract model: 68 P bytes per GS iteration (four passes) x it/s per GPU / peak; "
                               "the implementation moves 36 P (two passes)",
        "kernel_share_of_step": {k: v / tot for k, v in share.items()},
        # second roofline of the same kernels: the FP32 pipe.  Packed FADD2 / FMUL2 take 2 clk per warp and
        # SMSP, FFMA2 with three register operands 3 clk (tools/micro/pipe_overlap.cu); per thread (16 points) the fused column
        # kernel issues 384 FADD2 (2.07 clk) + 104 complex multiplies (FMUL2 + FFMA2: 4.43 clk per pair) + 58 + 48 constant
        # twiddle FMUL2 / FFMA2 (2.07 / 2.26) + ~150 scalar FMA-pipe instructions = ~1650 clk, the row kernel ~1600
        "fp32_pipe": (lambda clk: {
            "what": "FP32-pipe time of the launch (SASS instruction counts x measured issue cost, all SMs busy) / measured time",
            "lane_clk_per_thread": clk,
            "floor_ms": {k: (P / 16 / 32) * clk[k] / (148 * 4) / (clocks.get("sm_mhz", 1965.0) * 1e3) for k in clk},
            "frac": {k: (P / 16 / 32) * clk[k] / (148 * 4) / (clocks.get("sm_mhz", 1965.0) * 1e3) / kern[k]["avg_ms"]
                     for k in clk if k in kern},
        })({"col_fused": 1650.0, "row_fused": 1600.0}),
        "measured": "CUDA events around every launch of the same K steps, repeated right after the timed region",
        "kernels": kern,
    }

    # ---- CPU baseline + parity (the unmodified reference's NumPy path, bounded sample) ------------
    cpu_iters = 2
    cpu_value, cpu_dt, cpu_kind, cpu_amp = cpu_it_per_s(1, cpu_iters)
    cpu = {"value": cpu_value, "unit": "it/s", "cores": 1, "kind": cpu_kind,
           "sample": f"{cpu_iters} GS iterations at 4096^2, slm 4096^2 (+ trailing _populate_results transform), "
                     + ("unmodified reference (baseline/_ref), NumPy backend" if cpu_kind == "reference" else "oracle port")
                     + f", 1 process ({os.cpu_count()} host cores present), {cpu_dt:.1f} s"}
    parity = None
    if parity_amp is not None and cpu_amp is not None:
        err = float(np.linalg.norm(parity_amp.astype(np.float64) - cpu_amp) / np.linalg.norm(cpu_amp.astype(np.float64)))
        parity = {"rel_rmse_amp_ff": err, "iters": cpu_iters, "tolerance": 1e-5, "against": cpu_kind,
                  "what": "far-field amplitude after optimize(GS, maxiter=2) on the headline inputs, this library vs the CPU "
                          "baseline run of the same line"}

    line = {
        "metric": METRIC, "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "method": METHOD, "iters_per_step": ITERS, "shape": list(SHAPE),
                   "slm_shape": list(SHAPE), "parallelism": f"replicas x{world}",
                   "target": f"dense random: every far-field column tile is processed ({head_info[1]}/{head_info[2]} "
                             f"active, sparse path used: {bool(head_info[0])})",
                   "l2": "working set 201 MB/iteration (field 134 MB + weights 67 MB) > 126 MB L2, no flush needed",
                   "final_allgather_ms": ag_ms, "geometry": geometry,
                   "host_cores_bound": (f"{len(numa_cpus)} cores next to the GPU (sysfs local_cpulist)" if numa_cpus
                                        else "not bound")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": int(8 * P), "d2h_bytes_per_step": int(4 * P),
                "ms_per_step": e2e_ms / args.steps,
                "how": "C-ABI calls with pinned host buffers (target + phase up, phase down); two holograms in flight so "
                       "copies overlap kernels"},
        "gpu_launches": total_launches,
        "wall_ms_per_step": wall_ms / args.steps,
        "host_submit_ms_per_step": host_idle_queue_ms,
        "host_submit_note": "wall time of the host calls of one step (restore, reset, optimize: ~105 launches) issued into an "
                            "empty launch queue; host_blocked_ms_per_step = the same inside the K-step timed region, where the "
                            "host runs ahead until the queue is full and then blocks in the launch calls",
        "host_blocked_ms_per_step": 1e3 * host_s / args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    line.update(extras)
    emit(line)
    if cm is not None:
        cm.barrier()
        cm.close()


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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="headline only (skip the other configs)")
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
