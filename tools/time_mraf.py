#!/usr/bin/env python
"""Device time of MRAF + WGS (pixel feedback) at 1024^2 through the public API, sparse path on and off."""
import os, sys, ctypes as C, numpy as np
sys.path.insert(0, os.getcwd())
from slmsuite_b200 import Hologram, _lib
lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
rng = np.random.default_rng(0)
for sparse in (True, False):
    t = np.zeros((1024, 1024), np.float32)
    pts = rng.integers(300, 700, (2, 50)); t[pts[1], pts[0]] = 1
    t[100:200, 100:900] = np.nan
    h = Hologram(t, phase=rng.uniform(-3, 3, (1024, 1024)).astype(np.float32)); h.set_sparse(sparse)
    for m in ("WGS-Leonardo", "WGS-Kim"):
        h.optimize(m, maxiter=20, verbose=False)
        ms = C.c_float(); lib.slmgs_sync(h._ctx); lib.slmgs_timer_start(h._ctx)
        for _ in range(5): h.optimize(m, maxiter=20, verbose=False)
        lib.slmgs_timer_stop(h._ctx, C.byref(ms))
        print(f"MRAF + {m} 1024^2 20 it: {ms.value/5:.3f} ms -> {20/(ms.value/5)*1e3:.0f} it/s", h.sparse_info())
