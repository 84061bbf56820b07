"""
ONE CompressedSpotHologram spread over several ranks (pixel slabs, one all-reduce of the N spot accumulators per
iteration): world_size 2 and 3 over gloo on the emulation library against the single-process hologram and the oracle.
The GPU version of the same check is tools/sharded_compressed_demo.py (two B200s over NCCL, run with gpurun --gpus 2).
"""
import os
import socket
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libslmgs_emu.so")


def problem():
    rng = np.random.default_rng(21)
    slm = (45, 64)  # 45 rows: uneven slabs for 2 and 3 ranks
    yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
    grid = ((xx - slm[1] / 2) * 12.6, (yy - slm[0] / 2) * 12.6)
    v = np.vstack([rng.uniform(-0.03, 0.03, (2, 14)), rng.uniform(-2e-4, 2e-4, (1, 14))])
    spot_amp = rng.uniform(0.5, 1.5, 14)
    spot_amp[3] = np.nan
    spot_amp[9] = 0.0
    src = np.exp(-((xx - 30) ** 2 + (yy - 20) ** 2) / 900.0)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    args = dict(basis="kxy", spot_amp=spot_amp, slm_grid=grid, zernike_scaling=1.0 / 500.0, amp=src, phase=phase)
    opt = dict(method="WGS-Kim", maxiter=6, verbose=False, fix_phase_iteration=3)
    return v, args, opt


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from _gloo_comm import GlooComm

    from slmsuite_b200 import _lib
    from slmsuite_b200.compressed import ShardedCompressedSpotHologram

    _lib.use_library(EMU_LIB)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    v, args, opt = problem()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = ShardedCompressedSpotHologram(v, comm=GlooComm(), **{k: (np.array(x, copy=True) if isinstance(x, np.ndarray) else x)
                                                                 for k, x in args.items()})
        seen = []
        h.optimize(callback=(lambda holo: seen.append(holo.amp_ff.copy()) and False) if rank >= 0 else None, **opt)
    q.put((rank, h.phase, h.amp_ff, h.weights, h.farfield, bool(h.flags["fixed_phase"]), len(seen), tuple(h.slm_shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pixel_sharded_compressed_hologram_gloo(world, emu):
    import torch.multiprocessing as mp

    from oracle import compressed_oracle
    from slmsuite_b200 import CompressedSpotHologram

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        r = q.get(timeout=240)
        results[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v, args, opt = problem()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = CompressedSpotHologram(v, **{k: (np.array(x, copy=True) if isinstance(x, np.ndarray) else x)
                                           for k, x in args.items()})
        one.optimize(**opt)
        ref = compressed_oracle.OracleCompressedSpotHologram(v, **args)
        ref.optimize(**opt)

    def rel(a, b):
        a = np.asarray(a, dtype=np.complex128)
        b = np.asarray(b, dtype=np.complex128)
        m = ~(np.isnan(a) & np.isnan(b))
        return np.linalg.norm((a - b)[m]) / max(np.linalg.norm(b[m]), 1e-30)

    for rank in range(world):
        phase, amp_ff, weights, farfield, fixed, n_seen, shape = results[rank]
        assert shape == (45, 64) and phase.shape == (45, 64) and n_seen == opt["maxiter"]
        assert fixed == bool(one.flags["fixed_phase"]) == bool(ref.flags["fixed_phase"])
        # against the single-process device path: same arithmetic, different summation order of the pixel sum
        assert rel(amp_ff, one.amp_ff) <= 2e-6 and rel(weights, one.weights) <= 2e-6
        d = np.angle(np.exp(1j * (phase.astype(np.float64) - one.phase)))
        assert np.sqrt(np.mean(d ** 2)) <= 2e-5
        # against the oracle
        assert rel(amp_ff, ref.amp_ff) <= 1e-5 and rel(weights, ref.weights) <= 1e-5
        assert rel(farfield, ref.farfield) <= 1e-4
        # every rank holds the same full phase
        assert np.array_equal(phase, results[0][0])


@pytest.mark.gpu
def test_pixel_sharded_compressed_hologram_nccl(cuda):
    """The same over NCCL on two GPUs (skipped on a single-GPU box): tools/sharded_compressed_demo.py asserts the
    sharded result against the single-GPU hologram."""
    import glob
    import subprocess

    if len(glob.glob("/dev/nvidia[0-9]*")) < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tools", "sharded_compressed_demo.py"), "200"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "speed-up" in out.stdout
