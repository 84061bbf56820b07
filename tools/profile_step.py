#!/usr/bin/env python
"""Small driver for ncu: one hologram, a few iterations of the fused loop (used by profiles/README.md recipes).

    python tools/profile_step.py [--method WGS-Kim] [--shape 4096] [--slm 1152 1920] [--iters 12] [--dense]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import Hologram, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--method", default="WGS-Kim")
ap.add_argument("--shape", type=int, default=4096)
ap.add_argument("--slm", type=int, nargs=2, default=None)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--dense", action="store_true", help="dense random target instead of 64 spots")
ap.add_argument("--fix", type=int, default=4, help="fix_phase_iteration for WGS-Kim")
ap.add_argument("--lib", default=None, help="alternative build of libslmgs.so")
a = ap.parse_args()

_lib.use_library(a.lib or _lib.DEFAULT_LIBRARY)
shape = (a.shape, a.shape)
slm = tuple(a.slm) if a.slm else shape
rng = np.random.default_rng(1)
if a.dense:
    target = rng.random(shape, dtype=np.float32)
else:
    target = np.zeros(shape, dtype=np.float32)
    pts = rng.integers(0, a.shape, (2, 64))
    target[pts[1], pts[0]] = 1
phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
h = Hologram(target, phase=phase, slm_shape=slm)
kw = {"fix_phase_iteration": a.fix} if a.method == "WGS-Kim" else {}
h.optimize(a.method, maxiter=a.iters, verbose=False, **kw)
geo = np.zeros(4, dtype=np.int32)
h._lib.slmgs_launch_geometry(h._ctx, _lib.iptr(geo))
print("done", a.method, shape, slm, "launches", h._lib.slmgs_launch_count(h._ctx), "geometry", geo.tolist(),
      "amp_ff max", float(h.amp_ff.max()))
