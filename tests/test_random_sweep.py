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
