#!/usr/bin/env python
"""A/B of the 8192-point team column kernels (ColKernelT8, SLMGS_TEAMS8=1) against the plain 8192 kernels (SLMGS_TEAMS8=0):
results (far-field amplitude, phase, weights) and time per optimize(), dense 8192^2 fields.  One process per setting."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "GS": dict(method="GS", maxiter=6),
    "WGS-Kim": dict(method="WGS-Kim", maxiter=6, fix_phase_iteration=3),
    "WGS-Leonardo": dict(method="WGS-Leonardo", maxiter=4),
    "WGS-Nogrette": dict(method="WGS-Nogrette", maxiter=3),
}


def worker(out):
    import ctypes as C

    from slmsuite_b200 import Hologram, SpotHologram, _lib

    lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
    n = 8192
    rng = np.random.default_rng(1)
    target = rng.uniform(0.0, 1.0, (n, n)).astype(np.float32)
    phase = rng.uniform(-np.pi, np.pi, (n, n)).astype(np.float32)
    res = {}
    # a well-conditioned target for the power-law updates (a dense random target has far-field amplitudes next to zero,
    # where w *= (t / f)^p amplifies rounding differences between two FFT algorithms without bound)
    spots = np.zeros((n, n), dtype=np.float32)
    spots[rng.integers(0, n, 2000), rng.integers(0, n, 2000)] = 1.0
    for name, kw in CASES.items():
        h = Hologram(spots if name in ("WGS-Kim", "WGS-Leonardo") else target, phase=phase.copy(), slm_shape=(n, n))
        h.optimize(verbose=False, **kw)
        res[name + "/amp_ff"] = h.amp_ff[::7, ::5].copy()
        res[name + "/phase"] = h.phase[::7, ::5].copy()
        res[name + "/weights"] = np.asarray(h.weights)[::7, ::5].copy()
        if name in ("WGS-Kim", "WGS-Leonardo"):
            res[name + "/amp_ff@spots"] = h.amp_ff[spots > 0].copy()
            res[name + "/weights@spots"] = np.asarray(h.weights)[spots > 0].copy()
        t = C.c_float()
        lib.slmgs_sync(h._ctx)
        lib.slmgs_timer_start(h._ctx)
        h.optimize(verbose=False, **kw)
        lib.slmgs_timer_stop(h._ctx, C.byref(t))
        print(f"{name:14s} {kw['maxiter']} it: {t.value:8.2f} ms  ({t.value / kw['maxiter'] * 1e3:7.0f} us / it)", flush=True)
        del h
    v = np.random.default_rng(5).uniform(64, n - 64, (2, 10000))
    h = SpotHologram((n, n), v, basis="knm")
    h.reset_phase(phase)
    h.optimize("WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
    res["spot/amp_ff"] = h.amp_ff[::7, ::5].copy()
    res["spot/weights"] = np.asarray(h.weights)[::7, ::5].copy()
    t = C.c_float()
    lib.slmgs_sync(h._ctx)
    lib.slmgs_timer_start(h._ctx)
    h.optimize("WGS-Leonardo", maxiter=10, verbose=False, feedback="computational_spot")
    lib.slmgs_timer_stop(h._ctx, C.byref(t))
    print(f"spot feedback  10 it: {t.value:8.2f} ms  ({t.value / 10 * 1e3:7.0f} us / it)", flush=True)
    np.savez(out, **res)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "worker":
        worker(sys.argv[2])
        sys.exit(0)
    outs = {}
    for flag in ("0", "1"):
        out = f"/tmp/ab_teams8_{flag}.npz"
        print(f"== SLMGS_TEAMS8={flag}", flush=True)
        subprocess.run([sys.executable, __file__, "worker", out], check=True, env=dict(os.environ, SLMGS_TEAMS8=flag, SLMGS_SPARSE="0"))
        outs[flag] = np.load(out)
    print("== teams8 vs plain (rel-RMSE; phase: rms of the wrapped difference)")
    for k in outs["0"].files:
        a, b = outs["1"][k].astype(np.float64), outs["0"][k].astype(np.float64)
        if k.endswith("phase"):
            d = np.angle(np.exp(1j * (a - b)))
            print(f"{k:24s} {np.sqrt(np.mean(d * d)):.3e} rad")
        else:
            print(f"{k:24s} {np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30):.3e}")
