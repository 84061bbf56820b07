#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown of one optimize() on a 4096^2 hologram (library profile counters).
    python tools/profile_kernels.py [method] [slm_h slm_w] [iters]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import Hologram, _lib  # noqa: E402

lib = _lib.use_library(os.environ.get("SLMGS_LIB") or _lib.DEFAULT_LIBRARY)
method = sys.argv[1] if len(sys.argv) > 1 else "WGS-Kim"
slm = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4096, 4096)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rng = np.random.default_rng(0)
os.environ["SLMGS_SPARSE"] = "0"
h = Hologram(rng.random((4096, 4096), dtype=np.float32) + 0.05, phase=rng.uniform(-3, 3, slm).astype(np.float32), slm_shape=slm)
h.optimize(method, maxiter=iters, verbose=False)
lib.slmgs_profile_enable(h._ctx, 1)
h.optimize(method, maxiter=iters, verbose=False)
ms = (C.c_float * 6)()
cnt = (C.c_int * 6)()
lib.slmgs_profile_read(h._ctx, ms, cnt)
lib.slmgs_profile_enable(h._ctx, 0)
names = ["row_first", "row_fused", "row_last", "col_forward", "col_fused", "col_inverse"]
print(method, slm, iters, "iterations")
for k in range(6):
    if cnt[k]:
        print(f"  {names[k]:12s} {cnt[k]:3d} launches, {ms[k]/cnt[k]*1e3:8.1f} us each")
t = C.c_float()
lib.slmgs_sync(h._ctx)
lib.slmgs_timer_start(h._ctx)
h.optimize(method, maxiter=iters, verbose=False)
lib.slmgs_timer_stop(h._ctx, C.byref(t))
print(f"  optimize: {t.value:.3f} ms -> {iters / t.value * 1e3:.0f} it/s")
