"""Host-side logic of the drop-in classes: constructor invariants, error behaviour, flags and the
WGS-Kim state machine (restating the assertions of the reference's tests/holography/test_algorithms.py
and the error branches of _hologram.py).  Runs on the host emulation (CPU)."""
import warnings

import numpy as np
import pytest


def _spots(shape, n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.zeros(shape, dtype=np.float32)
    t[rng.integers(0, shape[0], n), rng.integers(0, shape[1], n)] = 1
    return t


def test_hologram_construction(emu):
    # reference tests/holography/test_algorithms.py:21-49
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(0)
    slm_shape, shape = (256, 256), (512, 512)
    phase_in = rng.uniform(-np.pi, np.pi, slm_shape).astype(np.float32)
    amp_in = (1 + rng.random(slm_shape)).astype(np.float32)
    h = Hologram(_spots(shape, 5), amp=amp_in, phase=phase_in, slm_shape=slm_shape)
    assert h.shape == shape and h.slm_shape == slm_shape
    assert h.dtype == np.float32 and h.dtype_complex == np.complex64
    d = h.get_phase() - phase_in
    assert np.allclose(d, d.flat[0]) and np.isclose(d.flat[0], np.pi)
    assert np.allclose(h.extract_phase(), h.get_phase())
    r = h.get_amp() / amp_in
    assert np.allclose(r, r.flat[0], rtol=1e-6)
    assert np.isclose(np.sqrt(np.sum(np.square(h.get_amp()))), 1, rtol=1e-6)
    assert h.amp_ff is None and h.phase_ff is None and h.iter == 0
    assert np.isclose(np.sqrt(np.nansum(np.square(h.target))), 1, rtol=1e-6)
    assert np.allclose(h.weights, h.target)


def test_scalar_amp_default(emu):
    from slmsuite_b200 import Hologram

    h = Hologram((64, 64), slm_shape=(20, 30))
    assert np.isscalar(h.amp) and np.isclose(h.amp, 1 / np.sqrt(600))
    assert h.phase.shape == (20, 30)
    assert np.all(np.abs(h.phase) <= np.pi)


def test_constructor_errors(emu):
    from slmsuite_b200 import Hologram

    t = _spots((64, 64), 3)
    with pytest.raises(ValueError, match="shape of the initial phase|shape of amplitude|shape of SLM"):
        Hologram(t, phase=np.zeros((32, 32), np.float32), slm_shape=(16, 16))
    with pytest.raises(ValueError, match="not supported"):
        Hologram(t, dtype=np.float16)
    with pytest.raises(ValueError, match="float32/complex64 only"):
        Hologram(t, dtype=np.float64)
    with pytest.raises(ValueError, match="powers of two"):
        Hologram(np.zeros((48, 64), np.float32))
    with pytest.raises(ValueError, match="too small"):
        Hologram(t, slm_shape=(128, 128))
    with pytest.raises(ValueError, match="propagation kernel"):
        Hologram(t, slm_shape=(32, 32), propagation_kernel=np.zeros((8, 8), np.float32))
    with pytest.raises(ValueError, match="Unexpected target"):
        Hologram(np.zeros((2, 3, 4), np.float32))
    h = Hologram(t)
    with pytest.raises(ValueError, match="not of slm_shape"):
        h.reset_phase(np.zeros((3, 3), np.float32))
    with pytest.raises(ValueError, match="do not match target shape"):
        h.set_weights(np.zeros((3, 3), np.float32))


def test_optimize_argument_errors(emu):
    # _hologram.py:1375-1410
    from slmsuite_b200 import Hologram

    h = Hologram(_spots((64, 64), 3))
    with pytest.raises(ValueError, match="Unrecognized method"):
        h.optimize("nope", maxiter=1, verbose=False)
    with pytest.raises(ValueError, match="Statistics group"):
        h.optimize("GS", maxiter=1, verbose=False, stat_groups=["bogus"])
    with pytest.raises(ValueError, match="Feedback 'bogus'"):
        h.optimize("GS", maxiter=1, verbose=False, feedback="bogus")
    with pytest.raises(ValueError, match="Must track statistics"):
        h.optimize("WGS-Kim", maxiter=3, verbose=False, fix_phase_efficiency=0.5)


def test_padded_shape():
    # _hologram.py:713-723
    from slmsuite_b200 import Hologram

    assert Hologram.get_padded_shape((1152, 1920)) == (2048, 2048)
    assert Hologram.get_padded_shape((1152, 1920), padding_order=2) == (4096, 4096)
    assert Hologram.get_padded_shape((720, 1280), square_padding=False) == (1024, 2048)
    assert Hologram.get_padded_shape((100, 100), padding_order=0) == (100, 100)


def test_flags_defaults_and_kim_history(emu):
    from slmsuite_b200 import Hologram

    h = Hologram(_spots((64, 64), 10), phase=np.zeros((64, 64), np.float32))
    h.optimize("WGS-Kim", maxiter=14, verbose=False)
    assert h.flags["feedback_exponent"] == 0.8 and h.flags["fix_phase_iteration"] == 10
    assert h.flags["feedback"] == "computational" and h.flags["fixed_phase"] is True
    hist = h.stats["flags"]["fixed_phase"]
    # recorded False for iterations 0..9, flips during iteration 9
    assert hist == [False] * 10 + [True] * 4
    assert h.stats["method"] == ["WGS-Kim"] * 14 and h.iter == 14
    # fixed_phase persists into a later GS call (reference quirk), non-Kim WGS resets it
    h.optimize("GS", maxiter=2, verbose=False)
    assert h.flags["fixed_phase"] is True
    h.optimize("WGS-Leonardo", maxiter=2, verbose=False)
    assert h.flags["fixed_phase"] is False
    assert h.iter == 18 and len(h.stats["method"]) == 18


def test_callback_contract(emu):
    # _hologram.py:1473-1475: called after the forward transform, truthy return breaks with phase untouched
    from slmsuite_b200 import Hologram

    phase0 = np.random.default_rng(1).uniform(-3, 3, (64, 64)).astype(np.float32)
    h = Hologram(_spots((64, 64), 6), phase=phase0)
    seen = []

    def cb(holo):
        seen.append((holo.iter, float(np.sum(np.square(holo.amp_ff)))))
        return holo.iter == 3

    h.optimize("WGS-Leonardo", maxiter=10, verbose=False, callback=cb)
    assert [s[0] for s in seen] == [0, 1, 2, 3] and h.iter == 3
    assert all(abs(s[1] - 1) < 1e-5 for s in seen)  # Parseval: ortho transform of a unit-norm near field
    h2 = Hologram(_spots((64, 64), 6), phase=phase0)
    h2.optimize("WGS-Leonardo", maxiter=3, verbose=False)
    assert np.allclose(h.phase, h2.phase, atol=2e-5)


def test_gs_validity_single_delta_gives_blaze(emu):
    # reference tests/holography/test_algorithms.py:51-84, without the 8-bit SLM quantisation
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(5)
    for method in ("GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette"):
        t = np.zeros((64, 64), dtype=np.float32)
        ky, kx = int(rng.integers(0, 64)), int(rng.integers(0, 64))
        t[ky, kx] = 1
        h = Hologram(t, phase=rng.uniform(-np.pi, np.pi, (64, 64)).astype(np.float32))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h.optimize(method, maxiter=20, verbose=False, stat_groups=["computational"])
        y, x = np.mgrid[0:64, 0:64]
        blaze = 2 * np.pi * ((kx - 32) * (x - 32) / 64.0 + (ky - 32) * (y - 32) / 64.0)
        err = np.angle(np.exp(1j * (h.get_phase() - blaze)))
        err = np.angle(np.exp(1j * (err - err.flat[0])))
        assert np.allclose(err, 0, atol=0.1), method
        assert h.stats["stats"]["computational"]["efficiency"][-1] > 0.99


def test_gs_convergence(emu):
    # reference tests/holography/test_algorithms.py:86-119
    from slmsuite_b200 import Hologram

    for method in ("GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette"):
        h = Hologram(_spots((64, 64), 20, seed=3), phase=np.random.default_rng(4).uniform(-3, 3, (64, 64)).astype(np.float32))
        h.optimize(method, maxiter=20, verbose=False, stat_groups=["computational"])
        st = h.stats["stats"]["computational"]
        assert st["efficiency"][-1] >= st["efficiency"][0]
        assert np.std(st["efficiency"][-5:]) < 0.05
        if "WGS" in method:
            assert st["std_err"][-1] <= st["std_err"][1]


def test_spot_hologram_construction_and_errors(emu):
    from slmsuite_b200 import SpotHologram

    h = SpotHologram.make_rectangular_array((64, 64), array_shape=(4, 3), array_pitch=(8, 10), basis="knm")
    assert h.spot_knm.shape == (2, 12) and h.spot_integration_width_knm == 3
    assert np.allclose(h.spot_amp, 1 / np.sqrt(12))
    assert np.count_nonzero(h.target) == 12
    assert np.isclose(np.sqrt(np.sum(np.square(h.target))), 1, rtol=1e-6)
    with pytest.raises(ValueError, match="outside SLM computational space"):
        SpotHologram((64, 64), np.array([[10.0, 70.0], [10.0, 10.0]]), basis="knm")
    with pytest.raises(ValueError, match="same length"):
        SpotHologram((64, 64), np.array([[10.0, 20.0], [10.0, 10.0]]), basis="knm", spot_amp=[1, 2, 3])
    with pytest.raises(AssertionError):
        SpotHologram((64, 64), np.array([[0.01], [0.01]]), basis="kxy")
    one = SpotHologram((64, 64), (20, 30), basis="knm")
    assert one.spot_integration_width_knm == 3 and one.spot_knm.shape == (2, 1)


def test_state_roundtrips_and_repeated_optimize(emu):
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(9)
    t = _spots((64, 128), 7, seed=2)
    h = Hologram(t, phase=rng.uniform(-3, 3, (40, 100)).astype(np.float32), slm_shape=(40, 100))
    w = rng.random((64, 128)).astype(np.float32)
    h.set_weights(w)
    assert np.array_equal(h.get_weights(), w)          # centred <-> rolled tile-major layout round trip
    h.weights = w * 2
    assert np.array_equal(h.weights, w * 2)
    pf = rng.uniform(-3, 3, (64, 128)).astype(np.float32)
    h.phase_ff = pf
    assert np.array_equal(h.phase_ff, pf)
    h.phase_ff = None
    assert h.phase_ff is None
    h.reset_weights()
    assert np.array_equal(h.weights, np.nan_to_num(h.target, nan=0))
    # maxiter=0 still populates the far field of the current phase (_hologram.py:1493)
    h.optimize("GS", maxiter=0, verbose=False)
    assert h.iter == 0 and h.amp_ff is not None and abs(float(np.sum(np.square(h.amp_ff))) - 1) < 1e-5
    # two optimize calls of 5 iterations == one call of 10 (state lives on the device between calls)
    a = Hologram(t, phase=np.zeros((40, 100), np.float32) + 0.3, slm_shape=(40, 100))
    b = Hologram(t, phase=np.zeros((40, 100), np.float32) + 0.3, slm_shape=(40, 100))
    a.optimize("WGS-Leonardo", maxiter=5, verbose=False)
    a.optimize("WGS-Leonardo", maxiter=5, verbose=False)
    b.optimize("WGS-Leonardo", maxiter=10, verbose=False)
    assert a.iter == b.iter == 10
    assert np.allclose(a.phase, b.phase, atol=3e-5) and np.allclose(a.weights, b.weights, rtol=1e-4, atol=1e-8)
    # reset() restores weights / iteration count / statistics, keeps the flags
    a.reset(reset_phase=False)
    assert a.iter == 0 and a.stats == {"method": [], "flags": {}, "stats": {}} and a.amp_ff is None
    assert a.flags["method"] == "WGS-Leonardo"
    assert np.array_equal(a.weights, np.nan_to_num(a.target, nan=0))


def test_set_target_renormalises_and_farfield_accessor(emu):
    from slmsuite_b200 import Hologram

    h = Hologram(_spots((64, 64), 4), phase=np.zeros((64, 64), np.float32))
    t2 = -3.0 * _spots((64, 64), 9, seed=5)                       # abs + L2 normalisation, _hologram.py:760-766
    h.set_target(t2, reset_weights=True)
    assert np.all(h.target >= 0) and np.isclose(np.sqrt(np.sum(np.square(h.target))), 1, rtol=1e-6)
    assert np.array_equal(h.weights, h.target)
    ff = h.get_farfield()
    assert ff.dtype == np.complex64 and ff.shape == (64, 64)
    # uniform amplitude, zero phase -> all the power in the centre pixel of the centred far field
    assert np.isclose(abs(ff[32, 32]), 1, rtol=1e-5) and np.isclose(np.sum(np.abs(ff) ** 2), 1, rtol=1e-5)
    assert np.allclose(h.amp_ff, np.abs(ff), atol=1e-7)
    # other padded shape / other depth (reference get_farfield(shape=, propagation_kernel=), _hologram.py:853-931)
    from oracle import gs_oracle

    rng = np.random.default_rng(12)
    ph = rng.uniform(-3, 3, (40, 56)).astype(np.float32)
    amp = (1 + rng.random((40, 56))).astype(np.float32)
    kern = rng.uniform(-2, 2, (40, 56)).astype(np.float32)
    g = Hologram((64, 64), amp=amp, phase=ph, slm_shape=(40, 56))
    o = gs_oracle.OracleHologram((128, 256), amp=amp, phase=ph, slm_shape=(40, 56), propagation_kernel=kern)
    o._forward()
    got = g.get_farfield(shape=(128, 256), propagation_kernel=kern)
    assert got.shape == (128, 256) and np.linalg.norm(got - o.farfield) / np.linalg.norm(o.farfield) < 1e-5
    o2 = gs_oracle.OracleHologram((128, 128), amp=amp, phase=ph, slm_shape=(40, 56))
    o2._forward()
    assert np.linalg.norm(g.get_farfield(shape=(128, 128)) - o2.farfield) / np.linalg.norm(o2.farfield) < 1e-5
    from scipy.ndimage import affine_transform

    aff = {"M": np.array([[1.1, 0.05], [-0.02, 0.9]]), "b": np.array([3.0, -2.0])}
    want = affine_transform(o2.farfield, aff["M"], offset=aff["b"], output_shape=(128, 128), order=3, mode="constant", cval=0)
    got = g.get_farfield(shape=(128, 128), affine=aff)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-4
    with pytest.raises(ValueError, match="powers of two"):
        g.get_farfield(shape=(100, 128))
    with pytest.raises(ValueError, match="does not match hologram shape"):
        h.set_target(np.zeros((32, 32), np.float32))


def test_spot_window_edges_and_feedback_names(emu):
    from slmsuite_b200 import SpotHologram

    # a spot on the border: the 3x3 integration window wraps like NumPy negative indices (analysis.take, clip=False)
    h = SpotHologram((64, 64), np.array([[0.0, 30.0, 40.0], [0.0, 20.0, 50.0]]), basis="knm")
    h.reset_phase(np.random.default_rng(3).uniform(-3, 3, (64, 64)).astype(np.float32))
    h.optimize("WGS-Leonardo", maxiter=5, verbose=False, feedback="computational_spot")
    assert h.iter == 5 and np.isfinite(h.weights).all()
    # a window that would run past the end of the array is an index error in the reference; here ValueError
    h2 = SpotHologram((64, 64), np.array([[63.0, 30.0], [63.0, 20.0]]), basis="knm")
    h2.reset_phase(np.zeros((64, 64), np.float32))
    h2.optimize("GS", maxiter=1, verbose=False)
    with pytest.raises(ValueError, match="out of bounds"):
        h2._window_power(h2.spot_knm, 3)
    with pytest.raises(IndexError, match="out of bounds"):   # the reference's analysis.take raises IndexError too
        h2.optimize("WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
    with pytest.raises(NotImplementedError, match="camera"):
        h.optimize("WGS-Leonardo", maxiter=2, verbose=False, feedback="experimental_spot")


def test_parse_cpulist_and_bind_is_harmless_without_a_gpu():
    """bind_host_to_device: sysfs cpulist parsing; without a visible GPU topology nothing is changed."""
    import os

    from slmsuite_b200 import _lib

    assert _lib._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _lib._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    out = _lib.bind_host_to_device(0)
    assert out is None or set(out) <= before
    os.sched_setaffinity(0, before)
