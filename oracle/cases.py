"""
Deterministic parity cases shared by the golden-vector generator, the oracle tests
and the GPU parity tests  --  TEST INFRASTRUCTURE ONLY.

Each case is a function ``build(Hologram, SpotHologram) -> (hologram, optimize_kwargs)``.
The three implementations (unmodified reference, ``oracle.gs_oracle`` restatement,
``slmsuite_b200`` product) expose the same constructor / ``optimize`` signatures
(SURVEY.md §8b), so one builder drives all three.  Inputs never rely on any
implementation's RNG: phases are always passed explicitly (reference RNG differs per
backend, slmsuite/holography/algorithms/_hologram.py:529-534).
"""

import numpy as np

PI = np.pi


def _phase(seed, shape):
    return np.random.default_rng(seed).uniform(-PI, PI, shape).astype(np.float32)


def _spots_target(seed, shape, n):
    """n unit pixels at seeded positions (form of reference tests/holography/test_algorithms.py:89-96)."""
    rng = np.random.default_rng(seed)
    t = np.zeros(shape, dtype=np.float32)
    for _ in range(n):
        t[rng.integers(0, shape[0]), rng.integers(0, shape[1])] = 1
    return t


def _gauss(shape, frac=0.35):
    y = np.linspace(-1, 1, shape[0])[:, None]
    x = np.linspace(-1, 1, shape[1])[None, :]
    return np.exp(-(x * x + y * y) / (2 * frac * frac)).astype(np.float32) + 0.01


CASES = {}


def case(name):
    def deco(fn):
        CASES[name] = fn
        return fn
    return deco


# ---- Hologram, slm_shape == shape -------------------------------------------------------
@case("gs_delta_64")
def _c(H, S):  # test_algorithms.py:51-84 (single far-field delta -> blaze)
    t = np.zeros((64, 64), dtype=np.float32)
    t[23, 41] = 1
    return H(target=t, phase=_phase(11, (64, 64))), dict(method="GS", maxiter=20)


@case("gs_dense_64")
def _c(H, S):
    t = np.random.default_rng(3).random((64, 64), dtype=np.float32)
    return H(target=t, phase=_phase(12, (64, 64))), dict(method="GS", maxiter=20)


def _mk_spots20(method, stats=True):
    def build(H, S):  # test_algorithms.py:86-119 (20 unit spots, 20 iterations, stats on)
        t = _spots_target(5, (64, 64), 20)
        kw = dict(method=method, maxiter=20)
        if stats:
            kw["stat_groups"] = ["computational"]
        return H(target=t, phase=_phase(13, (64, 64))), kw
    return build


for _m in ("GS", "WGS-Leonardo", "WGS-Kim", "WGS-Nogrette", "WGS-Wu", "WGS-tanh"):
    CASES["spots20_64_" + _m] = _mk_spots20(_m)
CASES["spots20_64_nostats_WGS-Leonardo"] = _mk_spots20("WGS-Leonardo", stats=False)
CASES["spots20_64_nostats_WGS-Kim"] = _mk_spots20("WGS-Kim", stats=False)


@case("rect_64x128_WGS-Leonardo")
def _c(H, S):
    t = _spots_target(6, (64, 128), 12)
    return H(target=t, phase=_phase(14, (64, 128))), dict(method="WGS-Leonardo", maxiter=15)


@case("rect_256x32_GS")
def _c(H, S):
    t = _spots_target(61, (256, 32), 9)
    return H(target=t, phase=_phase(141, (256, 32))), dict(method="GS", maxiter=12)


# ---- padded: slm_shape != shape, array amp, Kim long enough to fix the phase ------------
@case("padded_kim_128")
def _c(H, S):
    slm = (36, 60)
    t = _spots_target(7, (128, 128), 16)
    return (H(target=t, amp=_gauss(slm), phase=_phase(15, slm), slm_shape=slm),
            dict(method="WGS-Kim", maxiter=25))


@case("padded_odd_gs_128")
def _c(H, S):  # odd padding deltas: floor/ceil placement, toolbox/__init__.py:1701-1709
    slm = (37, 61)
    t = _spots_target(8, (128, 128), 10)
    return H(target=t, phase=_phase(16, slm), slm_shape=slm), dict(method="GS", maxiter=15)


@case("padded_dense_leonardo_1iter_128")
def _c(H, S):  # dense-target WGS is chaotic in fp32: pin 2 iterations only
    slm = (64, 64)
    t = np.random.default_rng(9).random((128, 128), dtype=np.float32) + 0.05
    return (H(target=t, phase=_phase(17, slm), slm_shape=slm),
            dict(method="WGS-Leonardo", maxiter=2))


@case("propagation_gs_64")
def _c(H, S):
    slm = (32, 32)
    y, x = np.mgrid[-1:1:32j, -1:1:32j]
    kern = (3.0 * (x * x + y * y)).astype(np.float32)
    t = _spots_target(10, (64, 64), 8)
    return (H(target=t, phase=_phase(18, slm), slm_shape=slm, propagation_kernel=kern),
            dict(method="GS", maxiter=12))


# ---- MRAF (NaN = noise region) -----------------------------------------------------------
def _mraf_target():
    t = np.full((64, 64), np.nan, dtype=np.float32)
    t[16:48, 16:48] = 0
    yy, xx = np.mgrid[0:64, 0:64]
    blob = np.exp(-((xx - 32.0) ** 2 + (yy - 30.0) ** 2) / 40.0).astype(np.float32)
    t[24:40, 24:40] = blob[24:40, 24:40]
    return t


@case("mraf_gs_64")
def _c(H, S):
    return H(target=_mraf_target(), phase=_phase(19, (64, 64))), dict(method="GS", maxiter=15)


@case("mraf_factor_leonardo_64")
def _c(H, S):
    return (H(target=_mraf_target(), phase=_phase(20, (64, 64))),
            dict(method="WGS-Leonardo", maxiter=6, mraf_factor=0.7))


@case("mraf_zero_factor_gs_64")
def _c(H, S):  # zero-region accumulator, _hologram.py:1511-1515, :1613-1616
    return (H(target=_mraf_target(), phase=_phase(27, (64, 64))),
            dict(method="GS", maxiter=10, zero_factor=0.5))


@case("mraf_zero_factor_kim_64")
def _c(H, S):
    return (H(target=_mraf_target(), phase=_phase(28, (64, 64))),
            dict(method="WGS-Kim", maxiter=9, zero_factor=0.25, mraf_factor=0.8, fix_phase_iteration=4))


# ---- SpotHologram -----------------------------------------------------------------------
@case("spot_rect_64_leonardo_spotfb")
def _c(H, S):
    h = S.make_rectangular_array((64, 64), array_shape=(4, 4), array_pitch=(8, 8), basis="knm")
    h.reset_phase(_phase(21, (64, 64)))
    return h, dict(method="WGS-Leonardo", maxiter=20, feedback="computational_spot",
                   stat_groups=["computational_spot"])


@case("spot_rect_padded_128_kim_spotfb")
def _c(H, S):
    h = S.make_rectangular_array((128, 128), array_shape=(5, 3), array_pitch=(12, 16), basis="knm",
                                 slm_shape=(48, 48))
    h.reset_phase(_phase(22, (48, 48)))
    return h, dict(method="WGS-Kim", maxiter=20, feedback="computational_spot",
                   stat_groups=["computational_spot"])


@case("spot_random_64_pixelfb")
def _c(H, S):
    v = np.random.default_rng(23).uniform(4, 60, (2, 12))
    amps = np.random.default_rng(24).uniform(0.5, 1.5, 12)
    h = S((64, 64), v, basis="knm", spot_amp=amps)
    h.reset_phase(_phase(25, (64, 64)))
    return h, dict(method="WGS-Leonardo", maxiter=15, feedback="computational")


@case("spot_null_mraf_64")
def _c(H, S):
    v = np.array([[20.0, 44.0, 32.0], [20.0, 24.0, 44.0]])
    nv = np.array([[32.0], [32.0]])
    h = S((64, 64), v, basis="knm", null_vectors=nv, null_radius=4)
    h.reset_phase(_phase(26, (64, 64)))
    return h, dict(method="WGS-Leonardo", maxiter=10, feedback="computational_spot")


# ---- MultiplaneHologram ---------------------------------------------------
# builders: (Hologram, SpotHologram, MultiplaneHologram) -> (parent, optimize kwargs)
MULTI_CASES = {}


def _multi(name):
    def deco(fn):
        MULTI_CASES[name] = fn
        return fn
    return deco


def _defocus(slm, strength):
    y, x = np.mgrid[-1:1:slm[0] * 1j, -1:1:slm[1] * 1j]
    return (strength * (x * x + y * y)).astype(np.float32)


@_multi("multi_two_planes_gs")
def _m(H, S, M):  # two depths of focus, same padded shape, array amplitude (the reference needs one)
    slm = (32, 48)
    ph = _phase(31, slm)
    a = H(_spots_target(32, (64, 64), 5), amp=_gauss(slm), phase=ph, slm_shape=slm)
    b = H(_spots_target(33, (64, 64), 7), amp=_gauss(slm), phase=ph, slm_shape=slm, propagation_kernel=_defocus(slm, 4.0))
    return M([a, b]), dict(method="GS", maxiter=12)


@_multi("multi_mixed_shapes_kim")
def _m(H, S, M):  # different padded shapes (different ortho scales), weights, WGS-Kim past the fixing iteration
    slm = (36, 60)
    ph = _phase(34, slm)
    amp = np.ones(slm, dtype=np.float32)
    a = H(_spots_target(35, (64, 64), 6), amp=amp, phase=ph, slm_shape=slm)
    b = H(_spots_target(36, (128, 128), 9), amp=amp, phase=ph, slm_shape=slm, propagation_kernel=_defocus(slm, -3.0))
    c = S.make_rectangular_array((64, 128), array_shape=(3, 2), array_pitch=(12, 10), basis="knm", amp=amp, slm_shape=slm)
    c.reset_phase(ph)
    return M([a, b, c], weights=[1, 2, 1.5]), dict(method="WGS-Kim", maxiter=16, fix_phase_iteration=6)


@_multi("multi_leonardo_stats")
def _m(H, S, M):
    slm = (64, 64)
    ph = _phase(37, slm)
    amp = _gauss(slm)
    a = H(_spots_target(38, (64, 64), 8), amp=amp, phase=ph)
    b = H(_spots_target(39, (64, 64), 8), amp=amp, phase=ph, propagation_kernel=_defocus(slm, 6.0))
    return M([a, b], weights=[3, 1]), dict(method="WGS-Leonardo", maxiter=10, stat_groups=["computational"])


def run_multi_case(name, Hologram, SpotHologram, MultiplaneHologram):
    parent, kw = MULTI_CASES[name](Hologram, SpotHologram, MultiplaneHologram)
    parent.optimize(verbose=False, **kw)
    return parent


def summarize_multi(parent):
    out = {"phase": np.asarray(parent.phase, dtype=np.float32), "iter": np.int64(parent.iter)}
    for i, h in enumerate(parent.holograms):
        out[f"child{i}/amp_ff"] = np.asarray(h.amp_ff, dtype=np.float32)
        out[f"child{i}/weights"] = np.asarray(h.weights, dtype=np.float32)
        out[f"child{i}/fixed_phase"] = np.int64(bool(h.flags.get("fixed_phase", False)))
        for group, d in h.stats["stats"].items():
            if "efficiency" in d:
                out[f"child{i}/stats/{group}/efficiency"] = np.asarray(d["efficiency"], dtype=np.float64)
    return out


def run_case(name, Hologram, SpotHologram):
    """Build, optimise (verbose off) and return the hologram."""
    holo, kw = CASES[name](Hologram, SpotHologram)
    holo.optimize(verbose=False, **kw)
    return holo


def summarize(holo):
    """The arrays/scalars a golden fixture records for one finished case."""
    out = {
        "phase": np.asarray(holo.phase, dtype=np.float32),
        "amp_ff": np.asarray(holo.amp_ff, dtype=np.float32),
        "weights": np.asarray(holo.weights, dtype=np.float32),
        "iter": np.int64(holo.iter),
        "fixed_phase": np.int64(bool(holo.flags.get("fixed_phase", False))),
    }
    for group, d in holo.stats["stats"].items():
        for key in ("efficiency", "uniformity", "pkpk_err", "std_err"):
            if key in d:
                out[f"stats/{group}/{key}"] = np.asarray(d[key], dtype=np.float64)
    return out
