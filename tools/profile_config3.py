import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, '/root/repo')
from slmsuite_b200 import SpotHologram, _lib
lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
h = SpotHologram.make_rectangular_array((4096, 4096), array_shape=(32, 32), array_pitch=(64, 64), basis="knm")
rng = np.random.default_rng(0)
h.reset_phase(rng.uniform(-3, 3, (4096, 4096)).astype(np.float32))
kw = dict(maxiter=100, verbose=False, feedback="computational_spot")
h.optimize("WGS-Leonardo", **kw)
ms = C.c_float()
lib.slmgs_sync(h._ctx); lib.slmgs_timer_start(h._ctx)
h.optimize("WGS-Leonardo", **kw)
lib.slmgs_timer_stop(h._ctx, C.byref(ms))
print("config3 100 it: %.2f ms -> %.0f it/s" % (ms.value, 100 / ms.value * 1e3), h.sparse_info())
