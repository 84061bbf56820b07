import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import ctypes as C
from slmsuite_b200 import Hologram, _lib
lib = _lib.use_library(_lib.DEFAULT_LIBRARY)
rng = np.random.default_rng(0)
for n, slm in ((512, (512, 512)), (512, (300, 200)), (1024, (1024, 1024)), (256, (256, 256))):
    target = rng.random((n, n), dtype=np.float32)
    phase = rng.uniform(-3, 3, slm).astype(np.float32)
    res = {}
    for loop in ("1", "0"):
        os.environ["SLMGS_LOOP"] = loop
        os.environ["SLMGS_SPARSE"] = "0"
        h = Hologram(target, phase=phase, slm_shape=slm)
        h.optimize("GS", maxiter=30, verbose=False)
        ph, aff = h.phase.copy(), h.amp_ff.copy()
        h.reset_phase(phase)
        h.optimize("GS", maxiter=30, verbose=False)
        ms = C.c_float()
        lib.slmgs_sync(h._ctx); lib.slmgs_timer_start(h._ctx)
        for _ in range(10):
            h.reset_phase(phase) if False else None
            h.optimize("GS", maxiter=30, verbose=False)
        lib.slmgs_timer_stop(h._ctx, C.byref(ms))
        res[loop] = (ph, aff, ms.value / 10, lib.slmgs_launch_count(h._ctx))
    same = res["1"][0].tobytes() == res["0"][0].tobytes() and res["1"][1].tobytes() == res["0"][1].tobytes()
    print(n, slm, "loop kernel %.3f ms/optimize (%.0f it/s)  plain %.3f ms (%.0f it/s)  bit-identical: %s  launches %d vs %d" % (
        res["1"][2], 30 / res["1"][2] * 1e3, res["0"][2], 30 / res["0"][2] * 1e3, same, res["1"][3], res["0"][3]))
