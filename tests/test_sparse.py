"""
Sparse far-field path of the fused loop (slmgs_set_sparse / prepare_sparse in slmgs_api.cu).

``farfield = weights * exp(i phase_ff)`` (_hologram.py:1601-1605) vanishes wherever the weights are zero, so the
column kernels may skip column tiles with all-zero weights and the row kernels need not move their columns.  The
claim is *identical results*: every case here is run twice -- automatic sparse and forced dense -- and compared
element-wise (plus against the NumPy oracle), and ``sparse_info()`` must confirm which path actually ran.
"""
import warnings

import numpy as np
import pytest

from oracle import gs_oracle


@pytest.fixture(autouse=True)
def _rebuild_populate(monkeypatch):
    """These tests compare a sparse run with a dense run BIT FOR BIT.  A sparse run always ends with
    `_populate_results` rebuilt from the stored phase (its row kernel only stored the active column tiles); a dense
    run normally skips that row pass (the last fused row kernel has already written the row spectrum of the final
    near field, equal to the rebuilt one within float rounding, ~1e-7).  SLMGS_POPULATE_REBUILD=1 makes the dense
    run rebuild too, so that the loop itself is what is compared."""
    monkeypatch.setenv("SLMGS_POPULATE_REBUILD", "1")


@pytest.fixture(autouse=True)
def _narrow_tiles(monkeypatch):
    """Four-column tiles on both backends (the emulation pretends to have 4 SMs and would otherwise pick tiles so
    wide that the small test problems have fewer than eight of them)."""
    monkeypatch.setenv("SLMGS_COL_THREADS", "32")


def _spot_target(shape, cols, rows_per_col, rng):
    t = np.zeros(shape, dtype=np.float32)
    for c in cols:
        t[rng.integers(0, shape[0], rows_per_col), c] = rng.uniform(0.5, 1.5, rows_per_col)
    return t


def _run(cls_kwargs, opt_kwargs, sparse, weights=None):
    from slmsuite_b200 import Hologram

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = Hologram(**cls_kwargs)
        h.set_sparse(sparse)
        if weights is not None:
            h.set_weights(weights)
        h.optimize(verbose=False, **opt_kwargs)
    return h


def _same(a, b, tol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= tol * scale, np.abs(a - b).max() / scale


CASES = {
    "gs": dict(method="GS", maxiter=6),
    "leonardo": dict(method="WGS-Leonardo", maxiter=6),
    "kim": dict(method="WGS-Kim", maxiter=8, fix_phase_iteration=3),
    "wu": dict(method="WGS-Wu", maxiter=5),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("padded", [False, True])
def test_sparse_equals_dense(case, padded, backend):
    rng = np.random.default_rng(3)
    shape = (128, 256)
    slm = (50, 120) if padded else shape
    target = _spot_target(shape, [3, 4, 77, 130, 255], 3, rng)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    kw = dict(target=target, phase=phase, slm_shape=slm)
    a = _run(kw, CASES[case], True)
    used, n_active, n_tiles = a.sparse_info()
    assert used and 0 < n_active <= 5 and n_tiles >= 8
    b = _run(kw, CASES[case], False)
    assert b.sparse_info()[0] is False
    # the active columns go through the same arithmetic in the same order; only the order of the atomic
    # accumulation of sum(w^2) may differ between two launches
    tol = 0.0 if case == "gs" else 2e-6
    _same(a.phase, b.phase, tol * 10)
    _same(a.amp_ff, b.amp_ff, tol)
    _same(a.weights, b.weights, tol)
    # and against the oracle
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = gs_oracle.OracleHologram(target, phase=phase, slm_shape=slm)
        ref.optimize(verbose=False, **CASES[case])
    err = np.linalg.norm(a.amp_ff - ref.amp_ff) / np.linalg.norm(ref.amp_ff)
    assert err <= 1e-5, err


def test_dense_target_stays_dense(backend):
    rng = np.random.default_rng(4)
    target = rng.random((64, 128), dtype=np.float32)
    phase = rng.uniform(-np.pi, np.pi, (64, 128)).astype(np.float32)
    h = _run(dict(target=target, phase=phase), dict(method="GS", maxiter=2), True)
    used, n_active, n_tiles = h.sparse_info()
    assert not used and n_active == n_tiles


def test_occupancy_follows_weights_and_target(backend):
    """set_weights / set_target / reset_weights invalidate the tile list."""
    rng = np.random.default_rng(5)
    shape = (64, 256)
    target = _spot_target(shape, [10, 200], 2, rng)
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    kw = dict(target=target, phase=phase)
    h = _run(kw, dict(method="GS", maxiter=2), True)
    used, n0, n_tiles = h.sparse_info()
    assert used and n0 <= 2
    # user weights on other columns: those tiles must become active and contribute
    w = np.zeros(shape, dtype=np.float32)
    w[5, 100] = 1.0
    w[20, 101] = 0.5
    w[40, 30] = 0.25
    a = _run(kw, dict(method="GS", maxiter=3), True, weights=w)
    b = _run(kw, dict(method="GS", maxiter=3), False, weights=w)
    assert a.sparse_info()[0]
    _same(a.phase, b.phase)
    _same(a.amp_ff, b.amp_ff)
    # a new dense target on the same object switches back to the dense loop
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a.set_target(rng.random(shape, dtype=np.float32), reset_weights=True)
        a.optimize("GS", maxiter=1, verbose=False)
    assert not a.sparse_info()[0]


def test_mraf_noise_region_tiles_are_active(backend):
    """MRAF: a NaN target (noise region) passes the field through (_hologram.py:1643-1653): its tiles stay active."""
    rng = np.random.default_rng(6)
    shape = (64, 256)
    target = _spot_target(shape, [20, 21, 150], 3, rng)
    target[10:30, 60:90] = np.nan
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    kw = dict(target=target, phase=phase)
    for opt in (dict(method="GS", maxiter=4), dict(method="GS", maxiter=4, mraf_factor=0.5)):
        a = _run(kw, opt, True)
        b = _run(kw, opt, False)
        used, n_active, n_tiles = a.sparse_info()
        assert used and n_active * 2 <= n_tiles and n_active >= 30 // max(256 // n_tiles, 1)
        _same(a.phase, b.phase)
        _same(a.amp_ff, b.amp_ff)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = gs_oracle.OracleHologram(target, phase=phase)
            ref.optimize(verbose=False, **opt)
        assert np.linalg.norm(a.amp_ff - ref.amp_ff) / np.linalg.norm(ref.amp_ff) <= 1e-5


@pytest.mark.parametrize("method", ["WGS-Leonardo", "WGS-Kim", "WGS-Nogrette"])
def test_spot_feedback_sparse_equals_dense(method, backend):
    """computational_spot feedback gathers w x w windows of |farfield| round every spot (_spots.py:1573-1624): the
    tiles under the windows are processed by the in-loop forward pass even where the weights are zero."""
    from slmsuite_b200 import SpotHologram

    rng = np.random.default_rng(7)
    shape = (128, 256)
    # spots on tile edges so that the 3-wide windows straddle neighbouring tiles
    xs = np.array([8, 15, 16, 64, 127, 200, 200, 201], dtype=float)
    ys = np.array([5, 40, 90, 64, 100, 20, 70, 110], dtype=float)
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    out = []
    for sparse in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h = SpotHologram(shape, np.vstack([xs, ys]), basis="knm")
            h.set_sparse(sparse)
            h.reset_phase(phase)
            h.optimize(method, maxiter=6, verbose=False, feedback="computational_spot", fix_phase_iteration=3)
        out.append(h)
    a, b = out
    assert a.sparse_info()[0] and not b.sparse_info()[0]
    _same(a.phase, b.phase, 2e-5)
    _same(a.weights, b.weights, 2e-6)
    _same(a.amp_ff, b.amp_ff, 2e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = gs_oracle.OracleSpotHologram(shape, np.vstack([xs, ys]), basis="knm")
        ref.reset_phase(phase)
        ref.optimize(method, maxiter=6, verbose=False, feedback="computational_spot", fix_phase_iteration=3)
    assert np.linalg.norm(a.amp_ff - ref.amp_ff) / np.linalg.norm(ref.amp_ff) <= 1e-5


def test_batch_has_per_hologram_occupancy(backend):
    from slmsuite_b200 import HologramBatch

    rng = np.random.default_rng(8)
    shape = (64, 256)
    targets = np.stack([_spot_target(shape, cols, 2, rng) for cols in ([5, 6], [100], [250, 30])])
    phases = rng.uniform(-np.pi, np.pi, (3,) + shape).astype(np.float32)
    res = []
    for sparse in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hb = HologramBatch(targets, phase=phases)
            hb.set_sparse(sparse)
            hb.optimize("WGS-Leonardo", maxiter=4, verbose=False)
        res.append(hb)
    a, b = res
    used, n_active, n_tiles = a.sparse_info()
    assert used and n_active <= 2  # the hologram with the most active tiles, not the union (5)
    _same(a.phase, b.phase, 2e-5)
    _same(a.amp_ff, b.amp_ff, 2e-6)
    _same(a.weights, b.weights, 2e-6)


def test_nogrette_sparse_equals_dense(backend):
    """WGS-Nogrette's mean(ratio) runs over the whole far field (_hologram.py:1851-1852); the ratio is exactly 1 where
    the target is zero, so tiles without target are skipped and counted.  Includes user weights on a column that has
    no target (inactive for the constraint's purposes only if its weights were zero) and a batch."""
    from slmsuite_b200 import HologramBatch

    rng = np.random.default_rng(9)
    shape = (128, 256)
    target = _spot_target(shape, [3, 4, 77, 130, 255], 3, rng)
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    kw = dict(target=target, phase=phase)
    opt = dict(method="WGS-Nogrette", maxiter=6)
    a = _run(kw, opt, True)
    b = _run(kw, opt, False)
    assert a.sparse_info()[0] and not b.sparse_info()[0]
    _same(a.phase, b.phase, 2e-5)
    _same(a.amp_ff, b.amp_ff, 2e-6)
    _same(a.weights, b.weights, 2e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = gs_oracle.OracleHologram(target, phase=phase)
        ref.optimize(verbose=False, **opt)
    assert np.linalg.norm(a.amp_ff - ref.amp_ff) / np.linalg.norm(ref.amp_ff) <= 1e-5
    assert np.linalg.norm(a.weights - ref.weights) / np.linalg.norm(ref.weights) <= 1e-5
    # weights on a target-free column
    w = target.copy()
    w[10, 200] = 0.3
    a = _run(kw, opt, True, weights=w)
    b = _run(kw, opt, False, weights=w)
    assert a.sparse_info()[0]
    _same(a.weights, b.weights, 2e-6)
    _same(a.amp_ff, b.amp_ff, 2e-6)
    # batch: every hologram counts its own skipped tiles
    targets = np.stack([_spot_target(shape, cols, 2, rng) for cols in ([5, 6, 7, 100], [100], [250, 30])])
    phases = rng.uniform(-np.pi, np.pi, (3,) + shape).astype(np.float32)
    res = []
    for sparse in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hb = HologramBatch(targets, phase=phases)
            hb.set_sparse(sparse)
            hb.optimize("WGS-Nogrette", maxiter=4, verbose=False)
        res.append(hb)
    assert res[0].sparse_info()[0]
    _same(res[0].weights, res[1].weights, 2e-6)
    _same(res[0].phase, res[1].phase, 2e-5)


@pytest.mark.parametrize("method,kw", [("WGS-Leonardo", {}), ("WGS-Kim", {"fix_phase_iteration": 2}),
                                       ("WGS-Leonardo", {"mraf_factor": 0.6}), ("WGS-tanh", {})])
def test_mraf_wgs_fused_sparse_equals_dense_and_oracle(method, kw, backend):
    """MRAF + WGS with pixel feedback runs in the fused loop (sum(w_new^2) pre-pass, immediate normalisation): the
    noise-region tiles and the tiles with weights are active, everything else is skipped."""
    rng = np.random.default_rng(12)
    shape = (64, 256)
    target = _spot_target(shape, [20, 21, 150, 151], 3, rng)
    target[10:30, 60:90] = np.nan
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    args = dict(target=target, phase=phase)
    opt = dict(method=method, maxiter=5, **kw)
    a = _run(args, opt, True)
    b = _run(args, opt, False)
    assert a.sparse_info()[0] and not b.sparse_info()[0]
    _same(a.phase, b.phase, 2e-5)
    _same(a.amp_ff, b.amp_ff, 2e-6)
    _same(np.nan_to_num(a.weights), np.nan_to_num(b.weights), 2e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = gs_oracle.OracleHologram(target, phase=phase)
        ref.optimize(verbose=False, **opt)
    assert np.linalg.norm(a.amp_ff - ref.amp_ff) / np.linalg.norm(ref.amp_ff) <= 1e-5
    assert np.linalg.norm(a.weights - ref.weights) / np.linalg.norm(ref.weights) <= 1e-5
    assert bool(a.flags["fixed_phase"]) == bool(ref.flags["fixed_phase"])


def test_reset_weights_restores_the_target_occupancy(backend):
    """reset_weights() sets weights = nan_to_num(target): the tile list falls back to the target's occupancy without a
    new device pass, and the next run equals a fresh hologram's."""
    rng = np.random.default_rng(13)
    shape = (64, 256)
    target = _spot_target(shape, [10, 200], 2, rng)
    phase = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    w = np.zeros(shape, dtype=np.float32)
    w[5, 100] = 1.0
    w[7, 130] = 1.0
    w[9, 60] = 1.0
    h = _run(dict(target=target, phase=phase), dict(method="WGS-Leonardo", maxiter=3), True, weights=w)
    n_custom = h.sparse_info()[1]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h.reset_phase(phase)
        h.reset(reset_phase=False)
        h.optimize("WGS-Leonardo", maxiter=3, verbose=False)
    fresh = _run(dict(target=target, phase=phase), dict(method="WGS-Leonardo", maxiter=3), True)
    assert h.sparse_info()[0] and h.sparse_info()[1] == fresh.sparse_info()[1] <= 2 < n_custom
    _same(h.phase, fresh.phase, 2e-5)
    _same(h.weights, fresh.weights, 2e-6)
