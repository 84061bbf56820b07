#!/usr/bin/env python
"""
ONE CompressedSpotHologram on N GPUs (pixel slabs + one all-reduce of the spot accumulators per iteration).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_compressed_demo.py

Checks the sharded result against the single-GPU hologram (rank 0) and reports time per iteration for both.
"""
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slmsuite_b200 import CompressedSpotHologram, _lib, comm as slm_comm  # noqa: E402
from slmsuite_b200.compressed import ShardedCompressedSpotHologram  # noqa: E402

rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
_lib.use_library(_lib.DEFAULT_LIBRARY)
cm = slm_comm.default()  # NCCL behind the C ABI (no torch): rendezvous on MASTER_ADDR / MASTER_PORT

n_spots = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
iters = 10
rng = np.random.default_rng(0)
slm = (1152, 1920)
yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
grid = ((xx - slm[1] / 2) * 12.6, (yy - slm[0] / 2) * 12.6)
v = np.vstack([rng.uniform(-0.03, 0.03, (2, n_spots)), rng.uniform(-1e-5, 1e-5, (1, n_spots))])
phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
args = dict(basis="kxy", slm_grid=grid, zernike_scaling=1.0 / 12000.0, phase=phase)
opt = dict(method="WGS-Kim", verbose=False, fix_phase_iteration=4)

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    h = ShardedCompressedSpotHologram(v, device=local, **args)
    h.optimize(maxiter=iters, **opt)  # warm-up + result for the check
    sharded_phase, sharded_amp = h.phase, h.amp_ff
    h.reset_phase(phase)
    h.reset(reset_phase=False)
    h.flags["fixed_phase"] = False
    h._check(h._lib.slmgs_comp_sync(h._ctx))
    cm.barrier()
    t0 = time.perf_counter()
    h.optimize(maxiter=iters, **opt)
    h._check(h._lib.slmgs_comp_sync(h._ctx))
    cm.barrier()
    t_sharded = (time.perf_counter() - t0) / iters

if rank == 0:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = CompressedSpotHologram(v, device=local, **args)
        one.optimize(maxiter=iters, **opt)
        one_phase, one_amp = one.phase, one.amp_ff
        one.reset_phase(phase)
        one.reset(reset_phase=False)
        one.flags["fixed_phase"] = False
        one._check(one._lib.slmgs_comp_sync(one._ctx))
        t0 = time.perf_counter()
        one.optimize(maxiter=iters, **opt)
        one._check(one._lib.slmgs_comp_sync(one._ctx))
        t_one = (time.perf_counter() - t0) / iters
    ea = float(np.linalg.norm(sharded_amp - one_amp) / np.linalg.norm(one_amp))
    d = np.angle(np.exp(1j * (sharded_phase.astype(np.float64) - one_phase)))
    print(f"sharded CompressedSpotHologram, {n_spots} spots (x, y, z), 1152x1920 SLM, WGS-Kim: {world} GPUs "
          f"{t_sharded*1e3:.3f} ms/iteration vs 1 GPU {t_one*1e3:.3f} ms/iteration -> speed-up {t_one/t_sharded:.2f}; "
          f"spot amplitude rel-RMSE vs single GPU {ea:.1e}, phase rms {np.sqrt(np.mean(d**2)):.1e} rad", flush=True)
    assert ea <= 2e-6 and np.sqrt(np.mean(d ** 2)) <= 5e-5
cm.barrier()
cm.close()
