"""HologramBatch (B holograms per launch) and the multi-rank sharding of a batch.
CPU: host emulation, world_size 2 (spawned processes), once over torch.distributed gloo (tests/_gloo_comm.py) and once
over the package's own torch-free TCP communicator (slmsuite_b200/comm.py).  The same single-rank checks run on the GPU
library when marked gpu; the NCCL all-gather behind the C ABI is exercised by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import EMU_LIB, ROOT


def _problem(B=3, shape=(64, 128), slm=(40, 100), seed=0):
    rng = np.random.default_rng(seed)
    T = np.zeros((B,) + shape, np.float32)
    for b in range(B):
        T[b, rng.integers(0, shape[0], 8), rng.integers(0, shape[1], 8)] = rng.uniform(0.5, 1, 8)
    P = rng.uniform(-np.pi, np.pi, (B,) + slm).astype(np.float32)
    return T, P, slm


def test_shard_bounds():
    from slmsuite_b200 import shard_bounds

    assert [shard_bounds(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert [shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    cover = []
    for r in range(3):
        lo, hi = shard_bounds(10, r, 3)
        cover += list(range(lo, hi))
    assert cover == list(range(10))


@pytest.mark.parametrize("method", ["GS", "WGS-Kim", "WGS-Nogrette"])
def test_batch_equals_independent_holograms(method, backend):
    from slmsuite_b200 import Hologram, HologramBatch

    T, P, slm = _problem()
    kw = dict(maxiter=12, verbose=False)
    if method == "WGS-Kim":
        kw["fix_phase_iteration"] = 5
    hb = HologramBatch(T, phase=P, slm_shape=slm)
    hb.optimize(method, **kw)
    assert hb.phase.shape == (3,) + slm and hb.amp_ff.shape == T.shape and len(hb) == 3
    for b in range(3):
        h = Hologram(T[b], phase=P[b], slm_shape=slm)
        h.optimize(method, **kw)
        # same kernels, same per-hologram arithmetic: only reduction order may differ
        assert np.allclose(hb.phase[b], h.phase, atol=1e-5)
        assert np.allclose(hb.amp_ff[b], h.amp_ff, atol=1e-7)
        assert np.allclose(hb.weights[b], h.weights, rtol=1e-5, atol=1e-9)


def test_batch_shared_target_and_stats(backend):
    from slmsuite_b200 import HologramBatch

    T, P, slm = _problem()
    hb = HologramBatch(T[0], phase=P, slm_shape=slm, batch=3)
    hb.optimize("WGS-Leonardo", maxiter=6, verbose=False, stat_groups=["computational"])
    eff = hb.stats["stats"]["computational"]["efficiency"]
    assert len(eff) == 6 and eff[-1].shape == (3,)
    assert np.all(eff[-1] > eff[0])
    with pytest.raises(ValueError, match="batch"):
        HologramBatch(T[0], phase=P, slm_shape=slm)
    with pytest.raises(ValueError, match="does not match"):
        HologramBatch(T, phase=P[:2], slm_shape=slm)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from _gloo_comm import GlooComm

    from slmsuite_b200 import _lib, optimize_sharded

    _lib.use_library(EMU_LIB)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, P, slm = _problem(B=5)
    phases, local = optimize_sharded(T, P, method="WGS-Leonardo", maxiter=8, slm_shape=slm, device=0, comm=GlooComm())
    q.put((rank, phases, None if local is None else len(local)))
    dist.barrier()
    dist.destroy_process_group()


def _worker_tcp(rank, world, port, q):
    """the package's own communicator: environment of a torchrun-style launcher, no torch"""
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    sys.path.insert(0, ROOT)
    from slmsuite_b200 import _lib, comm, optimize_sharded

    _lib.use_library(EMU_LIB)
    c = comm.default()
    assert (c.rank, c.world) == (rank, world) and not c.on_device()
    B = 3 if world == 4 else 5  # world 4, B 3: the last rank owns nothing and still joins the collective
    T, P, slm = _problem(B=B)
    phases, local = optimize_sharded(T, P, method="WGS-Leonardo", maxiter=8, slm_shape=slm, device=0)
    total = c.allreduce_host(np.array([float(rank + 1)]))
    c.barrier()
    q.put((rank, phases, None if local is None else len(local), float(total[0])))
    c.close()
    assert "torch" not in sys.modules


def test_two_rank_sharding_gloo(emu):
    """world_size 2 over gloo: ranks own holograms [0,3) and [3,5); after ONE all-gather both hold all phases,
    equal to the single-process batch."""
    import torch.multiprocessing as mp

    from slmsuite_b200 import HologramBatch

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(2):
        rank, phases, nlocal = q.get(timeout=240)
        results[rank] = (phases, nlocal)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == 3 and results[1][1] == 2
    T, P, slm = _problem(B=5)
    ref = HologramBatch(T, phase=P, slm_shape=slm)
    ref.optimize("WGS-Leonardo", maxiter=8, verbose=False)
    for rank in (0, 1):
        assert results[rank][0].shape == (5,) + slm
        assert np.allclose(results[rank][0], ref.phase, atol=1e-5)


@pytest.mark.parametrize("world", [2, 4])
def test_sharding_over_the_package_communicator(emu, world):
    """slmsuite_b200.comm (TCP rendezvous, no torch): contiguous shards, ONE all-gather, every rank holds every phase."""
    import multiprocessing as mp

    from slmsuite_b200 import HologramBatch

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_tcp, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, phases, nlocal, total = q.get(timeout=240)
        results[rank] = (phases, nlocal, total)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B = 3 if world == 4 else 5
    T, P, slm = _problem(B=B)
    ref = HologramBatch(T, phase=P, slm_shape=slm)
    ref.optimize("WGS-Leonardo", maxiter=8, verbose=False)
    if world == 4:
        assert [results[r][1] for r in range(4)] == [1, 1, 1, None]
    else:
        assert [results[r][1] for r in range(2)] == [3, 2]
    for rank in range(world):
        assert results[rank][0].shape == (B,) + slm
        assert np.allclose(results[rank][0], ref.phase, atol=1e-5)
        assert results[rank][2] == world * (world + 1) / 2
