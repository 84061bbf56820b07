#!/usr/bin/env python
"""Per-kernel breakdown of BASELINE configs[4] (SpotHologram, 10k spots, 8192^2, WGS-Leonardo, spot feedback)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import SpotHologram, _lib  # noqa: E402

lib = _lib.use_library(os.environ.get("SLMGS_LIB") or _lib.DEFAULT_LIBRARY)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
v = np.random.default_rng(5).uniform(64, n - 64, (2, 10000))
h = SpotHologram((n, n), v, basis="knm")
h.reset_phase(rng.uniform(-3, 3, (n, n)).astype(np.float32))
h.optimize("WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
lib.slmgs_profile_enable(h._ctx, 1)
h.optimize("WGS-Leonardo", maxiter=10, verbose=False, feedback="computational_spot")
ms = (C.c_float * 6)()
cnt = (C.c_int * 6)()
lib.slmgs_profile_read(h._ctx, ms, cnt)
lib.slmgs_profile_enable(h._ctx, 0)
names = ["row_first", "row_fused", "row_last", "col_forward", "col_fused", "col_inverse"]
for k in range(6):
    if cnt[k]:
        print(f"{names[k]:12s} {cnt[k]:3d} launches, {ms[k]/cnt[k]*1e3:8.1f} us each")
t = C.c_float()
lib.slmgs_sync(h._ctx)
lib.slmgs_timer_start(h._ctx)
h.optimize("WGS-Leonardo", maxiter=10, verbose=False, feedback="computational_spot")
lib.slmgs_timer_stop(h._ctx, C.byref(t))
print(f"10 iterations + populate: {t.value:.2f} ms", h.sparse_info())
