"""Phase -> SLM gray levels: oracle and product against golden vectors recorded from
the reference's unmodified ``SimulatedSLM.set_phase``.  Integer output: bit-exact."""
import os

import numpy as np
import pytest

from oracle import gray_oracle, make_golden_gray, ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return z["display"], int(z["bitdepth"])


@pytest.mark.parametrize("name", sorted(make_golden_gray.CASES))
def test_gray_oracle_matches_golden(name):
    shape, bitdepth, seed, corr = make_golden_gray.CASES[name]
    raw, correction = make_golden_gray.inputs(shape, seed, corr)
    display, bd = _gold(name)
    got = gray_oracle.phase2gray(raw + np.pi, bitdepth, correction)
    assert bd == bitdepth and got.dtype == display.dtype
    assert np.array_equal(got, display)


@pytest.mark.parametrize("name", sorted(make_golden_gray.CASES))
def test_product_gray_matches_golden(name, backend):
    from slmsuite_b200 import Hologram

    shape, bitdepth, seed, corr = make_golden_gray.CASES[name]
    raw, correction = make_golden_gray.inputs(shape, seed, corr)
    display, _ = _gold(name)
    h = Hologram((128, 256), phase=raw, slm_shape=shape)   # the device phase is `raw`; get_phase() adds pi
    got = h.get_phase_gray(bitdepth, phase_correction=correction)
    assert got.dtype == display.dtype and got.shape == display.shape
    assert np.array_equal(got, display)
    assert np.array_equal(got, gray_oracle.phase2gray(h.get_phase(), bitdepth, correction))


def test_gray_after_optimize_and_batch(backend):
    from slmsuite_b200 import Hologram, HologramBatch

    rng = np.random.default_rng(7)
    t = np.zeros((64, 64), np.float32)
    t[rng.integers(0, 64, 6), rng.integers(0, 64, 6)] = 1
    P = rng.uniform(-np.pi, np.pi, (2, 40, 56)).astype(np.float32)
    hb = HologramBatch(t, phase=P, slm_shape=(40, 56), batch=2)
    hb.optimize("WGS-Leonardo", maxiter=5, verbose=False)
    g = hb.get_phase_gray(8)
    assert g.shape == (2, 40, 56) and g.dtype == np.uint8
    for b in range(2):
        assert np.array_equal(g[b], gray_oracle.phase2gray(hb.get_phase()[b], 8))
    h = Hologram(t, phase=P[0], slm_shape=(40, 56))
    with pytest.raises(ValueError, match="bitdepth"):
        h.get_phase_gray(17)
    with pytest.raises(ValueError, match="phase_correction"):
        h.get_phase_gray(8, phase_correction=np.zeros((3, 3)))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_gray_accepted_by_reference_slm(emu):
    """The integer image is what the unmodified SLM class would have computed, and its set_phase accepts it."""
    import warnings

    from slmsuite_b200 import Hologram

    ref_loader.load_reference()
    from slmsuite.hardware.slms.simulated import SimulatedSLM

    raw = np.random.default_rng(8).uniform(-np.pi, np.pi, (48, 64)).astype(np.float32)
    h = Hologram((64, 64), phase=raw, slm_shape=(48, 64))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        slm = SimulatedSLM((64, 48))
        want = np.array(slm.set_phase(h.get_phase(), settle=False))
        got = h.get_phase_gray(slm.bitdepth)
        assert np.array_equal(got, want)
        assert np.array_equal(np.array(slm.set_phase(got, settle=False)), want)
