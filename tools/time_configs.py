#!/usr/bin/env python
"""Device-time per iteration of the BASELINE configs through the public API (CUDA events on the library stream)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import Hologram, HologramBatch, MultiplaneHologram, SpotHologram, _lib  # noqa: E402

lib = _lib.use_library(os.environ.get("SLMGS_LIB") or _lib.DEFAULT_LIBRARY)


def timed(h, reps, **kw):
    h.optimize(verbose=False, **kw)  # warm
    ms = C.c_float()
    lib.slmgs_sync(h._ctx)
    lib.slmgs_timer_start(h._ctx)
    for _ in range(reps):
        h.optimize(verbose=False, **kw)
    lib.slmgs_timer_stop(h._ctx, C.byref(ms))
    return ms.value / reps


def spots(shape, n, seed):
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, shape[0], (2, n))
    t = np.zeros(shape, dtype=np.float32)
    t[pts[1], pts[0]] = 1
    return t


def tag(h):
    used, n, m = h.sparse_info()
    return f"  [sparse far field: {n}/{m} column tiles]" if used else "  [dense]"


if "--dense" in sys.argv:  # force the dense loop (every column tile processed)
    sys.argv.remove("--dense")
    os.environ["SLMGS_SPARSE"] = "0"
which = sys.argv[1:] or ["1", "2", "2d", "3", "4", "5", "mp", "gray", "cam", "comp", "refbench"]
rng = np.random.default_rng(0)
if "1" in which:
    h = Hologram(rng.random((512, 512), dtype=np.float32), phase=rng.uniform(-3, 3, (512, 512)).astype(np.float32))
    ms = timed(h, 5, method="GS", maxiter=30)
    print(f"config1 GS 512^2 30 it: {ms:.3f} ms/optimize -> {30/ms*1e3:.0f} it/s" + tag(h))
if "2" in which:
    h = Hologram(spots((4096, 4096), 64, 1), phase=rng.uniform(-3, 3, (1152, 1920)).astype(np.float32), slm_shape=(1152, 1920))
    ms = timed(h, 3, method="WGS-Kim", maxiter=50)
    print(f"config2 WGS-Kim 4096^2 (slm 1152x1920) 50 it: {ms:.3f} ms/optimize -> {50/ms*1e3:.0f} it/s" + tag(h))
if "2d" in which:
    h = Hologram(rng.random((4096, 4096), dtype=np.float32), phase=rng.uniform(-3, 3, (4096, 4096)).astype(np.float32))
    ms = timed(h, 3, method="GS", maxiter=50)
    print(f"dense GS 4096^2 (slm 4096^2) 50 it: {ms:.3f} ms/optimize -> {50/ms*1e3:.0f} it/s  model frac {68*4096*4096*50/ms*1e3/6.538e12:.3f}")
    ms = timed(h, 3, method="WGS-Kim", maxiter=50)
    print(f"dense WGS-Kim 4096^2 50 it: {ms:.3f} ms/optimize -> {50/ms*1e3:.0f} it/s")
if "3" in which:
    h = SpotHologram.make_rectangular_array((4096, 4096), array_shape=(32, 32), array_pitch=(64, 64), basis="knm")
    h.reset_phase(rng.uniform(-3, 3, (4096, 4096)).astype(np.float32))
    ms = timed(h, 2, method="WGS-Leonardo", maxiter=100, feedback="computational_spot")
    print(f"config3 SpotHologram 32x32 on 4096^2 WGS-Leonardo spot feedback 100 it: {ms:.3f} ms/optimize -> {100/ms*1e3:.0f} it/s" + tag(h))
if "4" in which:
    B = 8
    T = np.stack([spots((2048, 2048), 100, 100 + b) for b in range(B)])
    P = rng.uniform(-3, 3, (B, 2048, 2048)).astype(np.float32)
    h = HologramBatch(T, phase=P)
    ms = timed(h, 3, method="GS", maxiter=50)
    print(f"config4 shard: batch of {B} 2048^2 GS 50 it: {ms:.3f} ms/optimize -> {B*50/ms*1e3:.0f} hologram-it/s per GPU" + tag(h))
if "5" in which:
    v = np.random.default_rng(5).uniform(64, 8192 - 64, (2, 10000))
    t0 = time.time()
    h = SpotHologram((8192, 8192), v, basis="knm")
    h.reset_phase(rng.uniform(-3, 3, (8192, 8192)).astype(np.float32))
    print(f"config5 ctor {time.time()-t0:.1f} s")
    ms = timed(h, 1, method="WGS-Leonardo", maxiter=20, feedback="computational_spot")
    print(f"config5 SpotHologram 10k spots 8192^2 WGS-Leonardo spot feedback 20 it: {ms:.3f} ms/optimize -> {20/ms*1e3:.0f} it/s" + tag(h))
if "mp" in which:
    slm = (1152, 1920)
    ph = rng.uniform(-3, 3, slm).astype(np.float32)
    amp = np.ones(slm, np.float32)
    yy, xx = np.mgrid[-1:1:slm[0] * 1j, -1:1:slm[1] * 1j]
    kids = [Hologram(spots((4096, 4096), 64, 10 + k), amp=amp, phase=ph, slm_shape=slm,
                     propagation_kernel=(float(k) * 5 * (xx * xx + yy * yy)).astype(np.float32)) for k in range(2)]
    m = MultiplaneHologram(kids)
    m.optimize("WGS-Kim", maxiter=5, verbose=False)
    lead = kids[0]
    ms = C.c_float()
    lib.slmgs_sync(lead._ctx)
    lib.slmgs_timer_start(lead._ctx)
    m.optimize("WGS-Kim", maxiter=20, verbose=False)
    lib.slmgs_timer_stop(lead._ctx, C.byref(ms))
    print(f"MultiplaneHologram 2 planes, 1152x1920 in 4096^2, WGS-Kim 20 it: {ms.value:.3f} ms -> {20/ms.value*1e3:.0f} it/s")
if "gray" in which:
    h = Hologram(spots((4096, 4096), 64, 1), phase=rng.uniform(-3, 3, (1152, 1920)).astype(np.float32), slm_shape=(1152, 1920))
    h.get_phase_gray(8)
    t0 = time.perf_counter()
    for _ in range(10):
        g = h.get_phase_gray(8)
    t1 = time.perf_counter()
    for _ in range(10):
        p = h.get_phase()
    t2 = time.perf_counter()
    print(f"get_phase_gray(8) 1152x1920: {(t1-t0)*100:.3f} ms per call ({g.nbytes/1e6:.1f} MB down) vs get_phase() {(t2-t1)*100:.3f} ms ({p.nbytes/1e6:.1f} MB down)")
if "cam" in which:
    from slmsuite_b200 import SimulatedCamera

    slm = (1024, 1024)
    yy, xx = np.mgrid[0:1200, 0:1600]
    knm = np.array([424.0 + yy * 1.0, 224.0 + xx * 1.0]) + 0.25
    cam = SimulatedCamera(slm, resolution=(1600, 1200), knm_cam=knm, shape_padded=(2048, 2048), bitdepth=8)
    cam.set_exposure(1e4)
    disp = rng.integers(0, 256, slm).astype(np.uint8)
    cam.get_image(disp, 256)
    t0 = time.perf_counter()
    for _ in range(10):
        img = cam.get_image(disp, 256)
    t1 = time.perf_counter()
    print(f"SimulatedCamera 1600x1200 of a 1024^2 SLM in 2048^2: {(t1-t0)*100:.3f} ms per get_image "
          f"(4 MB phase up, forward transform, sampling, {img.nbytes/1e6:.1f} MB image down)")
if "comp" in which:
    from slmsuite_b200 import CompressedSpotHologram

    slm = (1152, 1920)
    yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
    grid = ((xx - slm[1] / 2) * 12.6, (yy - slm[0] / 2) * 12.6)   # x / lambda for an 8 um pitch at 633 nm
    scaling = 1.0 / (2 * 6000.0)
    for dim, n_spots in ((2, 1000), (3, 1000), (2, 100)):
        v = rng.uniform(-0.03, 0.03, (dim, n_spots))
        if dim == 3:
            v[2] = rng.uniform(-1e-5, 1e-5, n_spots)
        h = CompressedSpotHologram(v, basis="kxy", slm_grid=grid, zernike_scaling=scaling,
                                   phase=rng.uniform(-3, 3, slm).astype(np.float32))
        h.optimize("WGS-Kim", maxiter=3, verbose=False)
        ms = C.c_float()
        lib.slmgs_comp_sync(h._ctx)
        lib.slmgs_comp_timer(h._ctx, 1, None)
        h.optimize("WGS-Kim", maxiter=10, verbose=False)
        lib.slmgs_comp_timer(h._ctx, 0, C.byref(ms))
        pairs = 2 * 10.5 * n_spots * slm[0] * slm[1]   # two maps per iteration (+ the trailing forward map)
        print(f"CompressedSpotHologram {n_spots} spots ({dim}-D) on a 1152x1920 SLM, WGS-Kim 10 it: {ms.value:.2f} ms -> "
              f"{10/ms.value*1e3:.1f} it/s, {pairs/ms.value/1e9:.2f} T (pixel, spot) pairs/s; "
              f"uniformity of |farfield|: {h.amp_ff.min()/h.amp_ff.max():.3f}")
if "refbench" in which:
    # the reference's own benchmark: tests/holography/test_algorithms.py:121-145 (1024^2, 20 spots, 20 iterations)
    for method in ("GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette"):
        h = Hologram(spots((1024, 1024), 20, 7), phase=rng.uniform(-3, 3, (1024, 1024)).astype(np.float32))
        ms = timed(h, 5, method=method, maxiter=20)
        print(f"reference benchmark test_gs_speed[{method}] 1024^2 20 it: {ms:.3f} ms/optimize -> {20/ms*1e3:.0f} it/s" + tag(h))
