"""
INTEGRATION.md section 2, executed: the reference-side binding (``integration/slmsuite_b200_binding.py``) subclasses the
LIVE, unmodified reference ``Hologram`` and overrides ``optimize_gs`` over the C ABI; results are compared with the
untouched reference run on the same inputs.

CPU (this container): reference from /root/reference through oracle/ref_loader.py, host-emulation build of the library.
GPU box: the reference as installed for bench.py's reference arm (baseline/_ref), libslmgs.so.
"""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))


def _reference_algorithms():
    from oracle import ref_loader

    if ref_loader.reference_available():
        return ref_loader.load_reference()
    from baseline import ref_arm

    if ref_arm.available():
        return ref_arm.load()
    pytest.skip("no reference tree (neither /root/reference nor baseline/_ref)")


def rel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _spots(shape, n, seed):
    rng = np.random.default_rng(seed)
    t = np.zeros(shape, dtype=np.float32)
    t[rng.integers(0, shape[0], n), rng.integers(0, shape[1], n)] = 1
    return t


CASES = [
    # (shape, slm_shape, method, maxiter, kwargs, target kind)
    ((128, 128), None, "GS", 12, {}, "dense"),
    ((256, 128), (200, 100), "GS", 8, {}, "spots"),
    ((128, 256), (100, 180), "WGS-Kim", 14, {"fix_phase_iteration": 5}, "spots"),
    ((128, 128), (96, 96), "WGS-Leonardo", 8, {}, "spots"),
    ((128, 128), (96, 96), "WGS-Nogrette", 6, {}, "spots"),
]


def _run_case(lib_path, case, seed):
    import slmsuite_b200_binding as binding

    alg = _reference_algorithms()
    shape, slm, method, maxiter, kw, kind = case
    rng = np.random.default_rng(seed)
    target = rng.random(shape, dtype=np.float32) if kind == "dense" else _spots(shape, 25, seed)
    slm = slm or shape
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    amp = (0.5 + rng.random(slm)).astype(np.float32)
    B200Hologram = binding.bind(alg.Hologram, lib_path)
    assert issubclass(B200Hologram, alg.Hologram)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = B200Hologram(target.copy(), amp=amp.copy(), phase=phase.copy(), slm_shape=slm)
        a.optimize(method, maxiter=maxiter, verbose=False, **kw)
        b = alg.Hologram(target.copy(), amp=amp.copy(), phase=phase.copy(), slm_shape=slm)
        b.optimize(method, maxiter=maxiter, verbose=False, **kw)
    assert a.iter == b.iter == maxiter
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert rel_rmse(np.abs(a.farfield), np.abs(b.farfield)) <= 1e-5
    assert bool(a.flags.get("fixed_phase", False)) == bool(b.flags.get("fixed_phase", False))
    assert a.stats["flags"]["fixed_phase"] == b.stats["flags"]["fixed_phase"]
    nf = np.abs(b.nearfield)
    i0, i1 = (shape[0] - slm[0]) // 2, shape[0] - -(-(shape[0] - slm[0]) // 2)
    i2, i3 = (shape[1] - slm[1]) // 2, shape[1] - -(-(shape[1] - slm[1]) // 2)
    ok = nf[i0:i1, i2:i3] > 1e-4 * nf.max()
    d = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(d[ok] ** 2)) <= 2e-5
    # get_phase() is the reference's own accessor on the subclass
    assert np.allclose(a.get_phase(), a.phase + np.pi)
    # a second optimize() call continues from the state the first one left (weights, phase, Kim's frozen phase)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a.optimize(method, maxiter=3, verbose=False, **kw)
        b.optimize(method, maxiter=3, verbose=False, **kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 2e-5


@pytest.mark.parametrize("case", CASES, ids=[f"{c[2]}-{c[0][0]}x{c[0][1]}" for c in CASES])
def test_reference_subclass_over_c_abi_emu(emu_library, case):
    _run_case(emu_library, case, seed=11)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [((1024, 1024), (800, 600), "WGS-Kim", 12, {"fix_phase_iteration": 5}, "spots")],
                         ids=lambda c: f"{c[2]}-{c[0][0]}x{c[0][1]}")
def test_reference_subclass_over_c_abi_cuda(cuda_library, cuda, case):
    _run_case(cuda_library, case, seed=12)


def test_callback_falls_back_to_the_reference_loop(emu_library):
    import slmsuite_b200_binding as binding

    alg = _reference_algorithms()
    B200Hologram = binding.bind(alg.Hologram, emu_library)
    t = _spots((64, 64), 10, 3)
    seen = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h = B200Hologram(t, phase=np.zeros((64, 64), dtype=np.float32))
        h.optimize("GS", maxiter=4, verbose=False, callback=lambda holo: seen.append(holo.iter) or False)
    assert seen == [0, 1, 2, 3]
