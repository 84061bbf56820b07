"""
Parity of the product path (Python host classes -> C ABI -> kernels) with the reference.

Every case of ``oracle/cases.py`` is run through ``slmsuite_b200.Hologram`` / ``SpotHologram`` and
compared with the golden vectors recorded from the UNMODIFIED reference (tests/golden, made by
oracle/make_golden.py).  ``backend`` = "emu" runs the host emulation of the kernel sources (CPU
suite); ``backend`` = "cuda" (marked gpu) runs the sm_100a library on the B200.

Tolerances (north_star: far-field amplitude within 1e-5 rel-RMSE of the reference, fp32):
  amp_ff, weights : rel-RMSE <= 1e-5          phase : wrapped rms <= 2e-5 rad
  statistics      : rel <= 1e-4               dense-target WGS (chaotic in fp32): 2e-4
"""
import os
import warnings

import numpy as np
import pytest

from oracle import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LOOSE = {"padded_dense_leonardo_1iter_128": 20.0}


def rel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = ~(np.isnan(a) & np.isnan(b))
    return np.linalg.norm((a - b)[m]) / max(np.linalg.norm(b[m]), 1e-30)


def run_product(name):
    from slmsuite_b200 import Hologram, SpotHologram

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cases.run_case(name, Hologram, SpotHologram)


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_matches_reference_golden(name, backend):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        gold = {k: z[k] for k in z.files}
    holo = run_product(name)
    got = cases.summarize(holo)
    k = LOOSE.get(name, 1.0)
    assert set(got) == set(gold)
    assert int(got["iter"]) == int(gold["iter"])
    assert int(got["fixed_phase"]) == int(gold["fixed_phase"])
    assert rel_rmse(got["amp_ff"], gold["amp_ff"]) <= 1e-5 * k
    assert rel_rmse(got["weights"], gold["weights"]) <= 1e-5 * k
    dphi = np.angle(np.exp(1j * (got["phase"].astype(np.float64) - gold["phase"].astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5 * k
    for key in gold:
        if key.startswith("stats/"):
            assert got[key].shape == gold[key].shape
            assert rel_rmse(got[key], gold[key]) <= 1e-4 * k, key


@pytest.mark.parametrize("name", ["spots20_64_nostats_WGS-Leonardo", "spots20_64_nostats_WGS-Kim", "gs_dense_64",
                                  "padded_kim_128", "mraf_gs_64", "mraf_factor_leonardo_64", "spot_null_mraf_64",
                                  "spot_random_64_pixelfb"])
def test_fused_and_stepped_paths_agree(name, backend):
    """The same case through the fused two-kernel loop and through the stepped entry points
    (forced by a no-op callback) must agree: both implement _hologram.py:1465-1490."""
    from slmsuite_b200 import Hologram, SpotHologram

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fused = cases.summarize(cases.run_case(name, Hologram, SpotHologram))
        holo, kw = cases.CASES[name](Hologram, SpotHologram)
        holo.optimize(verbose=False, callback=lambda h: False, **kw)
    stepped = cases.summarize(holo)
    assert rel_rmse(stepped["amp_ff"], fused["amp_ff"]) <= 1e-5
    assert rel_rmse(stepped["weights"], fused["weights"]) <= 1e-5
    assert int(stepped["fixed_phase"]) == int(fused["fixed_phase"])


@pytest.mark.parametrize("method", ["WGS-Nogrette", "WGS-Wu", "WGS-tanh"])
def test_fused_sequence_other_weightings_vs_oracle(method, backend):
    """No statistics, no callback: the fused launch sequence (Nogrette: forward pass + update kernels + fused
    kernels; Wu / tanh: in-kernel update with fast math) against the oracle."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(31)
    target = np.zeros((64, 128), dtype=np.float32)
    target[rng.integers(0, 64, 15), rng.integers(0, 128, 15)] = rng.uniform(0.5, 1.0, 15)
    phase = rng.uniform(-np.pi, np.pi, (48, 100)).astype(np.float32)
    kw = dict(method=method, maxiter=12, verbose=False)
    a = Hologram(target, phase=phase, slm_shape=(48, 100))
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleHologram(target, phase=phase, slm_shape=(48, 100))
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5


def test_spot_feedback_fused_sequence_vs_oracle(backend):
    """SpotHologram with computational_spot feedback and no statistics runs the fused launch sequence."""
    from oracle import gs_oracle
    from slmsuite_b200 import SpotHologram

    phase = np.random.default_rng(41).uniform(-np.pi, np.pi, (128, 128)).astype(np.float32)
    kw = dict(method="WGS-Kim", maxiter=14, verbose=False, feedback="computational_spot", fix_phase_iteration=6)
    a = SpotHologram.make_rectangular_array((128, 128), array_shape=(5, 4), array_pitch=(10, 14), basis="knm")
    a.reset_phase(phase)
    a.optimize(**kw)
    b = gs_oracle.OracleSpotHologram.make_rectangular_array((128, 128), array_shape=(5, 4), array_pitch=(10, 14), basis="knm")
    b.reset_phase(phase)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert a.flags["fixed_phase"] == b.flags["fixed_phase"] is True


def test_kim_fix_phase_efficiency_vs_oracle(backend):
    """WGS-Kim fixing on an efficiency threshold (_hologram.py:1559-1572) needs per-iteration statistics: stepped path."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(51)
    target = np.zeros((64, 64), dtype=np.float32)
    target[rng.integers(0, 64, 12), rng.integers(0, 64, 12)] = 1
    phase = rng.uniform(-np.pi, np.pi, (64, 64)).astype(np.float32)
    kw = dict(method="WGS-Kim", maxiter=12, verbose=False, stat_groups=["computational"], fix_phase_efficiency=0.5,
              fix_phase_iteration=100)
    a = Hologram(target, phase=phase)
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleHologram(target, phase=phase)
        b.optimize(**kw)
    assert a.flags["fixed_phase"] == b.flags["fixed_phase"] is True
    assert a.stats["flags"]["fixed_phase"] == b.stats["flags"]["fixed_phase"]
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert rel_rmse(a.stats["stats"]["computational"]["efficiency"], b.stats["stats"]["computational"]["efficiency"]) <= 1e-4


def test_oracle_side_by_side_seeded(backend):
    """Product vs the oracle restatement on a case that is not in the golden set."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(77)
    target = np.zeros((128, 256), dtype=np.float32)
    target[rng.integers(0, 128, 30), rng.integers(0, 256, 30)] = rng.uniform(0.5, 1.0, 30)
    phase = rng.uniform(-np.pi, np.pi, (100, 180)).astype(np.float32)
    amp = (1 + 0.3 * rng.random((100, 180))).astype(np.float32)
    kw = dict(method="WGS-Kim", maxiter=14, verbose=False, fix_phase_iteration=5)
    a = Hologram(target, amp=amp, phase=phase, slm_shape=(100, 180))
    a.optimize(**kw)
    b = gs_oracle.OracleHologram(target, amp=amp, phase=phase, slm_shape=(100, 180))
    b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert a.flags["fixed_phase"] == b.flags["fixed_phase"]
    assert a.stats["flags"]["fixed_phase"] == b.stats["flags"]["fixed_phase"]


@pytest.mark.parametrize("method,kw", [("GS", {}), ("WGS-Leonardo", {}), ("WGS-Kim", {"fix_phase_iteration": 2}),
                                       ("WGS-Nogrette", {})])
def test_long_columns_vs_oracle(method, kw, backend):
    """Columns of 2048 points take the code paths reserved for long columns (L1 prefetch and private-slot staging of
    the constraint's image loads, slmgs_kernels.h): a tall 2048 x 16 problem against the oracle."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(11)
    shape, slm = (2048, 16), (700, 12)
    target = np.zeros(shape, dtype=np.float32)
    target[rng.integers(0, 2048, 40), rng.integers(0, 16, 40)] = rng.uniform(0.5, 1.5, 40)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = Hologram(target, phase=phase, slm_shape=slm)
        a.optimize(method, maxiter=4, verbose=False, **kw)
        b = gs_oracle.OracleHologram(target, phase=phase, slm_shape=slm)
        b.optimize(method, maxiter=4, verbose=False, **kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5
