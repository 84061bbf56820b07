"""
Seeded random sweep of small problems through the product path against the oracle: shapes (square, wide, tall),
SLM shapes of any parity (ragged zero padding), every method, spot / dense / MRAF targets, scalar / array amplitude,
propagation kernel, sparse and dense far-field paths.  Complements the golden cases with combinations nobody wrote down.

Tolerances as in tests/test_parity.py (amp_ff, weights 1e-5 rel-RMSE; phase 2e-5 rad rms); dense-target WGS is chaotic in
fp32, so those cases run a single weight update.
"""
import warnings

import numpy as np
import pytest

from oracle import gs_oracle

METHODS = ["GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette", "WGS-Wu", "WGS-tanh"]
SIZES = [16, 32, 64, 128, 256]


def rel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = ~(np.isnan(a) & np.isnan(b))
    return np.linalg.norm((a - b)[m]) / max(np.linalg.norm(b[m]), 1e-30)


def make_case(seed):
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.choice(SIZES)), int(rng.choice(SIZES))
    if rng.random() < 0.5:
        h, w = H, W
    else:
        h, w = int(rng.integers(max(1, H // 4), H + 1)), int(rng.integers(max(1, W // 4), W + 1))
    method = METHODS[int(rng.integers(len(METHODS)))]
    kind = ["spots", "spots", "dense", "mraf"][int(rng.integers(4))]
    target = np.zeros((H, W), dtype=np.float32)
    n = int(rng.integers(1, max(2, min(40, H * W // 8))))
    if kind == "dense":
        target = rng.random((H, W), dtype=np.float32)
    else:
        target[rng.integers(0, H, n), rng.integers(0, W, n)] = rng.uniform(0.2, 1.5, n).astype(np.float32)
        if kind == "mraf":
            y0, x0 = int(rng.integers(0, H // 2)), int(rng.integers(0, W // 2))
            blk = target[y0:y0 + H // 4 + 1, x0:x0 + W // 4 + 1]
            blk[blk == 0] = np.nan
    kw = {}
    if method == "WGS-Kim":
        kw["fix_phase_iteration"] = int(rng.integers(1, 4))
    if kind == "mraf" and rng.random() < 0.5:
        kw["mraf_factor"] = float(rng.uniform(0.2, 1.0))
    maxiter = int(rng.integers(1, 6))
    if kind == "dense" and method != "GS":
        maxiter = 2
    if kind == "mraf" and method != "GS":
        maxiter = min(maxiter, 3)
    amp = None
    if rng.random() < 0.4:
        yy, xx = np.mgrid[-1:1:h * 1j, -1:1:w * 1j] if h > 1 and w > 1 else (np.zeros((h, w)), np.zeros((h, w)))
        amp = np.exp(-(xx ** 2 + yy ** 2)).astype(np.float32)
    prop = rng.uniform(-1, 1, (h, w)).astype(np.float32) if rng.random() < 0.25 else None
    phase = rng.uniform(-np.pi, np.pi, (h, w)).astype(np.float32)
    return dict(target=target, slm=(h, w), method=method, kw=kw, maxiter=maxiter, amp=amp, prop=prop, phase=phase,
                kind=kind)


@pytest.mark.parametrize("seed", range(48))
def test_random_case_matches_oracle(seed, backend):
    from slmsuite_b200 import Hologram

    c = make_case(seed)
    args = dict(amp=None if c["amp"] is None else c["amp"].copy(), phase=c["phase"], slm_shape=c["slm"],
                propagation_kernel=c["prop"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = Hologram(c["target"], **args)
        a.optimize(c["method"], maxiter=c["maxiter"], verbose=False, **c["kw"])
        args["amp"] = None if c["amp"] is None else c["amp"].copy()
        b = gs_oracle.OracleHologram(c["target"], **args)
        b.optimize(c["method"], maxiter=c["maxiter"], verbose=False, **c["kw"])
    loose = 20.0 if (c["kind"] == "dense" and c["method"] != "GS") else 1.0
    info = (seed, c["target"].shape, c["slm"], c["method"], c["kind"], c["maxiter"], a.sparse_info())
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5 * loose, info
    assert rel_rmse(a.weights, b.weights) <= 1e-5 * loose, info
    assert bool(a.flags.get("fixed_phase", False)) == bool(b.flags.get("fixed_phase", False)), info
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    # a pixel whose near field is (numerically) zero has an arbitrary phase on both sides: compare where it is not
    nf = np.abs(b.nearfield)
    i0, i1, i2, i3 = gs_oracle.crop_bounds(b.shape, b.slm_shape)
    ok = nf[i0:i1, i2:i3] > 1e-4 * nf.max()
    assert np.sqrt(np.mean(dphi[ok] ** 2)) <= 2e-5 * loose, info


def make_spot_case(seed):
    rng = np.random.default_rng(5000 + seed)
    H, W = int(rng.choice([32, 64, 128, 256])), int(rng.choice([32, 64, 128, 256]))
    n = int(rng.integers(2, min(25, H - 12, W - 12)))
    # distinct integer-ish positions away from the border (the integration windows must fit)
    xs = rng.choice(np.arange(6, W - 6), size=n, replace=False).astype(float) + rng.uniform(-0.3, 0.3, n)
    ys = rng.choice(np.arange(6, H - 6), size=n, replace=False).astype(float) + rng.uniform(-0.3, 0.3, n)
    method = ["GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette", "WGS-Wu", "WGS-tanh"][int(rng.integers(6))]
    feedback = ["computational_spot", "computational"][int(rng.integers(2))]
    kw = {}
    if method == "WGS-Kim":
        kw["fix_phase_iteration"] = int(rng.integers(1, 4))
    ctor = {}
    if rng.random() < 0.3:
        k = int(rng.integers(1, 4))
        ctor["null_vectors"] = np.vstack([rng.uniform(8, W - 8, k), rng.uniform(8, H - 8, k)])
        ctor["null_radius"] = int(rng.integers(1, 4))
    if rng.random() < 0.5:
        ctor["spot_amp"] = rng.uniform(0.5, 1.5, n)
    padded = rng.random() < 0.4
    slm = (int(rng.integers(H // 2, H + 1)), int(rng.integers(W // 2, W + 1))) if padded else (H, W)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    return dict(shape=(H, W), v=np.vstack([xs, ys]), method=method, feedback=feedback, kw=kw, ctor=ctor, slm=slm,
                phase=phase, maxiter=int(rng.integers(2, 6)))


@pytest.mark.parametrize("seed", range(32))
def test_random_spot_case_matches_oracle(seed, backend):
    from slmsuite_b200 import SpotHologram

    c = make_spot_case(seed)
    out = []
    for cls in (SpotHologram, gs_oracle.OracleSpotHologram):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ctor = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in c["ctor"].items()}
            h = cls(c["shape"], c["v"].copy(), basis="knm", slm_shape=c["slm"], phase=c["phase"], **ctor)
            h.optimize(c["method"], maxiter=c["maxiter"], verbose=False, feedback=c["feedback"], **c["kw"])
        out.append(h)
    a, b = out
    info = (seed, c["shape"], c["slm"], c["method"], c["feedback"], sorted(c["ctor"]), c["maxiter"], a.sparse_info())
    assert a.spot_integration_width_knm == b.spot_integration_width_knm, info
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5, info
    assert rel_rmse(a.weights, b.weights) <= 1e-5, info
    assert bool(a.flags.get("fixed_phase", False)) == bool(b.flags.get("fixed_phase", False)), info
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5, info


@pytest.mark.parametrize("seed", range(16))
def test_random_compressed_case_matches_oracle(seed, backend):
    from oracle import compressed_oracle
    from slmsuite_b200 import CompressedSpotHologram

    rng = np.random.default_rng(9000 + seed)
    h, w = int(rng.integers(24, 97)), int(rng.integers(24, 97))
    yy, xx = np.mgrid[0:h, 0:w]
    grid = ((xx - w / 2) * 12.6, (yy - h / 2) * 12.6)
    scaling = 1.0 / float(rng.uniform(300, 900))
    n = int(rng.integers(2, 30))  # (one spot: the reference's set_target squeezes the vector to 0-d and raises, _spots.py:935-940)
    kind = int(rng.integers(4))
    if kind == 0:
        basis, v = "kxy", rng.uniform(-0.03, 0.03, (2, n))
    elif kind == 1:
        basis, v = "kxy", np.vstack([rng.uniform(-0.03, 0.03, (2, n)), rng.uniform(-2e-4, 2e-4, (1, n))])
    elif kind == 2:
        basis, v = "zernike", np.vstack([rng.uniform(-25, 25, (2, n)), rng.uniform(-3, 3, (2, n))])  # [2, 1, 4, 3]
    else:
        basis = [1, 2, 5, 7, 8, 9, 12]  # x and y anywhere in the list, coma, trefoil, spherical
        v = np.vstack([rng.uniform(-20, 20, (2, n)), rng.uniform(-1, 1, (5, n))])
    amp_n = rng.uniform(0.5, 1.5, n)
    if n >= 4 and rng.random() < 0.5:
        amp_n[int(rng.integers(n))] = np.nan
        amp_n[int(rng.integers(n))] = 0.0
        if not np.any(np.nan_to_num(amp_n) > 0):
            amp_n[0] = 1.0
    method = METHODS[int(rng.integers(len(METHODS)))]
    kw = {"fix_phase_iteration": int(rng.integers(1, 4))} if method == "WGS-Kim" else {}
    if np.any(np.isnan(amp_n)) and rng.random() < 0.5:
        kw["mraf_factor"] = float(rng.uniform(0.3, 1.0))
    src = np.exp(-((xx - w / 2) ** 2 + (yy - h / 2) ** 2) / (0.4 * h * w)) if rng.random() < 0.5 else None
    args = dict(basis=basis, spot_amp=amp_n, slm_grid=grid, zernike_scaling=scaling, amp=src,
                phase=rng.uniform(-np.pi, np.pi, (h, w)).astype(np.float32))
    maxiter = int(rng.integers(1, 6))
    out = []
    for cls in (CompressedSpotHologram, compressed_oracle.OracleCompressedSpotHologram):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = {k: (np.array(x, copy=True) if isinstance(x, np.ndarray) else x) for k, x in args.items()}
            o = cls(v.copy(), **a)
            o.optimize(method, maxiter=maxiter, verbose=False, **kw)
        out.append(o)
    a, b = out
    info = (seed, (h, w), n, basis, method, maxiter, kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5, info
    assert rel_rmse(a.weights, b.weights) <= 1e-5, info
    assert bool(a.flags.get("fixed_phase", False)) == bool(b.flags.get("fixed_phase", False)), info
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    nf = np.abs(b.nearfield)
    ok = nf > 1e-4 * nf.max()
    assert np.sqrt(np.mean(dphi[ok] ** 2)) <= 1e-4, info


@pytest.mark.parametrize("seed", range(12))
def test_consecutive_optimize_calls_match_oracle(seed, backend):
    """optimize() called twice (second call starts with iter > 0, i.e. with a weight update in its first iteration and
    whatever normalisation the first call left pending), with a method switch in between."""
    from slmsuite_b200 import Hologram

    c = make_case(100 + seed)
    if c["kind"] == "dense":
        c["target"] = np.where(np.random.default_rng(seed).random(c["target"].shape) < 0.02, c["target"], 0).astype(np.float32)
        c["target"][0, 0] = 1.0
    second = METHODS[(METHODS.index(c["method"]) + 1 + seed) % len(METHODS)]
    args = dict(phase=c["phase"], slm_shape=c["slm"], propagation_kernel=c["prop"])
    out = []
    for cls in (Hologram, gs_oracle.OracleHologram):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h = cls(c["target"], amp=None if c["amp"] is None else c["amp"].copy(), **args)
            h.optimize(c["method"], maxiter=2, verbose=False, **c["kw"])
            kw2 = {"fix_phase_iteration": 2} if second == "WGS-Kim" else {}
            h.optimize(second, maxiter=3, verbose=False, **kw2)
        out.append(h)
    a, b = out
    info = (seed, c["target"].shape, c["slm"], c["method"], second, c["kind"], a.sparse_info())
    assert int(a.iter) == int(b.iter) == 5
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5, info
    assert rel_rmse(a.weights, b.weights) <= 1e-5, info
    assert bool(a.flags.get("fixed_phase", False)) == bool(b.flags.get("fixed_phase", False)), info
