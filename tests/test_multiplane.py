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


@pytest.mark.parametrize("method,kw", [("GS", {}), ("WGS-Leonardo", {}), ("WGS-Kim", {"fix_phase_iteration": 3})])
def test_multiplane_fused_update_sparse_equals_dense_and_oracle(method, kw, backend, monkeypatch):
    """The fused child iteration (slmgs_run_accumulate) with the in-kernel update (sum(w_new^2) pre-pass, immediate
    normalisation) and the sparse far field: spot targets on wide far fields, against the dense loop and the oracle."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram, MultiplaneHologram

    monkeypatch.setenv("SLMGS_COL_THREADS", "32")  # four-column tiles on both backends
    rng = np.random.default_rng(31)
    slm = (40, 100)
    amp = np.exp(-np.linspace(-1, 1, slm[1]) ** 2)[None, :] * np.ones(slm)
    ph = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    yy, xx = np.mgrid[-1:1:slm[0] * 1j, -1:1:slm[1] * 1j]

    def spots(shape, cols, seed):
        r = np.random.default_rng(seed)
        t = np.zeros(shape, dtype=np.float32)
        for c in cols:
            t[r.integers(0, shape[0], 2), c] = r.uniform(0.5, 1.5, 2)
        return t

    def build(H, M):
        a = H(spots((64, 256), [5, 100, 101], 1), amp=amp.astype(np.float32), phase=ph, slm_shape=slm)
        b = H(spots((128, 256), [30, 200], 2), amp=amp.astype(np.float32), phase=ph, slm_shape=slm,
              propagation_kernel=(2.5 * (xx * xx + yy * yy)).astype(np.float32))
        return M([a, b], weights=[1.0, 0.7])

    res = []
    for sparse in (True, False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = build(Hologram, MultiplaneHologram)
            for h in m.holograms:
                h.set_sparse(sparse)
            m.optimize(method, maxiter=7, verbose=False, **kw)
        res.append(m)
    a, b = res
    assert all(h.sparse_info()[0] for h in a.holograms) and not any(h.sparse_info()[0] for h in b.holograms)
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = build(gs_oracle.OracleHologram, gs_oracle.OracleMultiplaneHologram)
        o.optimize(method, maxiter=7, verbose=False, **kw)
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - o.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5
    for ha, ho in zip(a.holograms, o.holograms):
        assert rel_rmse(ha.amp_ff, ho.amp_ff) <= 1e-5
        assert rel_rmse(ha.weights, ho.weights) <= 1e-5
