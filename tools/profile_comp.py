#!/usr/bin/env python
"""Small driver for ncu: a CompressedSpotHologram of 1000 (x, y) spots on a 1152x1920 SLM, three WGS-Kim iterations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import CompressedSpotHologram, _lib  # noqa: E402

_lib.use_library(_lib.DEFAULT_LIBRARY)
rng = np.random.default_rng(0)
slm = (1152, 1920)
yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
grid = ((xx - slm[1] / 2) * 12.6, (yy - slm[0] / 2) * 12.6)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
v = rng.uniform(-0.03, 0.03, (2, n))
h = CompressedSpotHologram(v, basis="kxy", slm_grid=grid, zernike_scaling=1.0 / 12000.0,
                           phase=rng.uniform(-3, 3, slm).astype(np.float32))
h.optimize("WGS-Kim", maxiter=3, verbose=False)
print("done", n, "spots, launches", h._lib.slmgs_comp_launch_count(h._ctx), "amp_ff", float(h.amp_ff.min()), float(h.amp_ff.max()))
