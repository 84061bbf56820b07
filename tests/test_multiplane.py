"""MultiplaneHologram against golden vectors recorded from the reference's unmodified
MultiplaneHologram (oracle/make_golden.py), on the host emulation (CPU) and on the B200 (gpu)."""
import os
import warnings

import numpy as np
import pytest

from oracle import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name", sorted(cases.MULTI_CASES))
def test_multiplane_matches_reference_golden(name, backend):
    from slmsuite_b200 import Hologram, MultiplaneHologram, SpotHologram

    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        gold = {k: z[k] for k in z.files}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = cases.summarize_multi(cases.run_multi_case(name, Hologram, SpotHologram, MultiplaneHologram))
    assert set(got) == set(gold)
    assert int(got["iter"]) == int(gold["iter"])
    dphi = np.angle(np.exp(1j * (got["phase"].astype(np.float64) - gold["phase"].astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5
    for k in gold:
        if k.endswith("amp_ff") or k.endswith("weights"):
            assert rel_rmse(got[k], gold[k]) <= 1e-5, k
        elif k.endswith("fixed_phase"):
            assert int(got[k]) == int(gold[k]), k
        elif "/stats/" in k:
            assert rel_rmse(got[k], gold[k]) <= 1e-4, k


def test_multiplane_api_and_errors(backend):
    from slmsuite_b200 import Hologram, MultiplaneHologram

    rng = np.random.default_rng(0)
    slm = (24, 40)
    ph = rng.uniform(-3, 3, slm).astype(np.float32)
    t = np.zeros((64, 64), np.float32)
    t[10, 20] = t[40, 50] = 1
    a = Hologram(t, phase=ph, slm_shape=slm)
    b = Hologram(t.T.copy(), phase=ph * 0, slm_shape=slm)
    m = MultiplaneHologram([a, b])
    assert len(m) == 2 and np.allclose(m.weights, 1 / np.sqrt(2)) and m.target is None and m.shape == slm
    assert np.array_equal(b.phase, a.phase) and np.array_equal(m.phase, ph)     # children share the first child's phase
    with pytest.raises(RuntimeError, match="set_target"):
        m.set_target(t)
    with pytest.raises(ValueError, match="recursion"):
        MultiplaneHologram([m])
    with pytest.raises(ValueError, match="child holograms"):
        MultiplaneHologram([a, "nope"])
    with pytest.raises(ValueError, match="slm_shape"):
        MultiplaneHologram([a, Hologram(t, slm_shape=(32, 32))])
    with pytest.raises(ValueError, match="Unrecognized method"):
        m.optimize("nope", maxiter=1, verbose=False)
    seen = []
    m.optimize("WGS-Leonardo", maxiter=6, verbose=False, callback=lambda p: seen.append(p.iter) or p.iter == 3)
    assert seen == [0, 1, 2, 3] and m.iter == 3 and a.iter == 3 and b.iter == 3
    assert a.flags["method"] == b.flags["method"] == "WGS-Leonardo"
    assert np.array_equal(a.phase, b.phase)
    g = m.get_phase_gray(8)
    assert g.shape == slm and g.dtype == np.uint8
    m.reset(reset_phase=False)
    assert m.iter == 0 and a.iter == 0 and a.amp_ff is None
