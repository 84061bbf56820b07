#!/usr/bin/env python
"""Diagnostic: per-phase SM-clock stamps of the fused kernels (needs slmsuite_b200/libslmgs_trace.so, built with
-DSLMGS_TRACE for N = 4096; see slmsuite_b200/csrc/slmgs_launch.h).  Prints, for a few blocks, the cycles spent in every
phase and behind every barrier of one launch on the dense 4096^2 GS loop."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slmsuite_b200 import Hologram, _lib  # noqa: E402

lib = _lib.use_library(os.path.join(ROOT, "slmsuite_b200", "libslmgs_trace.so"))
raw = C.CDLL(os.path.join(ROOT, "slmsuite_b200", "libslmgs_trace.so"))
method = sys.argv[1] if len(sys.argv) > 1 else "GS"
rng = np.random.default_rng(0)
shape = (4096, 4096)
h = Hologram(rng.random(shape, dtype=np.float32), phase=rng.uniform(-3, 3, shape).astype(np.float32))
h.optimize(method, maxiter=12, verbose=False)
for klass, name in ((11, "ColKernel fused"), (30, "ColKernelP (persistent, TMA)"), (21, "RowKernel fused")):
    lib.slmgs_sync(h._ctx)
    raw.slmgs_trace_enable(klass)
    h.optimize(method, maxiter=12, verbose=False)
    lib.slmgs_sync(h._ctx)
    raw.slmgs_trace_enable(0)
    buf = (C.c_longlong * (8 * 2 * 64))()
    raw.slmgs_trace_read(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(8, 2, 64)
    print(f"== {name}: cycles since block start; rows = (block, first/last thread); columns = stamps "
          "[start, (phase done, barrier passed) ...]")
    for b in range(3):
        for th in range(2):
            v = t[b, th]
            if v[0] == 0:
                continue
            n = int(np.max(np.nonzero(v)[0])) + 1
            rel = v[:n] - v[0]
            rel[v[:n] == 0] = -1
            print(f"  block {b} thread {'first' if th == 0 else 'last '}: " + " ".join(str(int(x)) for x in rel))
    t[:] = 0
