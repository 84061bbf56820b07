"""
CompressedSpotHologram ("next" row 4): ``slmsuite_b200.CompressedSpotHologram`` against results recorded
from the UNMODIFIED reference (tests/golden/compressed_*.npz, made by oracle/make_golden_compressed.py from
_spots.py:178-1019 on its NumPy backend).

Tolerances (north_star: far-field amplitude within 1e-5 rel-RMSE, fp32): spot amplitudes ``amp_ff`` and ``weights``
rel-RMSE <= 1e-5; near-field phase wrapped rms <= 1e-4 rad (the kernel phase spans hundreds of radians and the
reference accumulates it in float32; the oracle itself sits at 7e-6 from the reference).
"""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import compressed_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "compressed_*.npz")))
RUN = {
    "compressed_2d_leonardo": ("kxy", "WGS-Leonardo", 6, {}),
    "compressed_3d_kim": ("kxy", "WGS-Kim", 8, {"fix_phase_iteration": 3}),
    "compressed_zernike5_gs": ([2, 1, 4, 3, 5], "GS", 5, {}),
    "compressed_mraf_leonardo": ("kxy", "WGS-Leonardo", 6, {}),
    "compressed_2d_nogrette": ("kxy", "WGS-Nogrette", 5, {}),
}


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    m = ~(np.isnan(a) & np.isnan(b))
    return np.linalg.norm((a - b)[m]) / max(np.linalg.norm(b[m]), 1e-30)


def phase_rms(a, b):
    d = np.angle(np.exp(1j * (np.asarray(a, np.float64) - np.asarray(b, np.float64))))
    return float(np.sqrt(np.mean(d ** 2)))


def build(cls, g, name, **extra):
    basis = RUN[name][0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cls(g["spot_vectors"], basis=basis, spot_amp=g["spot_amp"].copy(), slm_grid=(g["x_grid"], g["y_grid"]),
                   zernike_scaling=float(g["zernike_scaling"]), amp=g["amp"], phase=g["phase0"], **extra)


def check(h, g):
    assert int(h.iter) == int(g["iter"])
    assert bool(h.flags.get("fixed_phase", False)) == bool(g["fixed_phase"])
    assert rel(h.amp_ff, g["amp_ff"]) <= 1e-5
    assert rel(h.weights, g["weights"]) <= 1e-5
    assert rel(np.abs(h.farfield), np.abs(g["farfield"])) <= 1e-5
    assert phase_rms(h.phase, g["phase"]) <= 1e-4
    # the complex spot amplitudes themselves (global phase included)
    assert rel(h.farfield, g["farfield"]) <= 1e-4


def test_golden_cases_exist():
    assert set(NAMES) == set(RUN)


def test_zernike_monomials_known_polynomials():
    from slmsuite_b200.compressed import monomial_table, zernike_monomials

    assert zernike_monomials(2) == {(1, 0): 1}                          # x
    assert zernike_monomials(1) == {(0, 1): 1}                          # y
    assert zernike_monomials(4) == {(2, 0): 2, (0, 2): 2, (0, 0): -1}   # 2 r^2 - 1
    assert zernike_monomials(3) == {(1, 1): 2}                          # 2 x y
    assert zernike_monomials(5) == {(2, 0): 1, (0, 2): -1}              # x^2 - y^2
    assert zernike_monomials(12) == {(4, 0): 6, (2, 2): 12, (0, 4): 6, (2, 0): -6, (0, 2): -6, (0, 0): 1}
    for idx in range(15):
        assert zernike_monomials(idx) == compressed_oracle.zernike_monomials(idx)
    px, py, c = monomial_table(np.array([2, 1, 4]))
    assert c.shape == (len(px), 3) and len(px) == 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    g = load(name)
    _basis, method, maxiter, kw = RUN[name]
    o = build(compressed_oracle.OracleCompressedSpotHologram, g, name)
    assert np.abs(o.spot_zernike - g["spot_zernike"]).max() < 1e-9
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o.optimize(method, maxiter=maxiter, verbose=False, **kw)
    check(o, g)


@pytest.mark.parametrize("name", NAMES)
def test_product_matches_reference(name, backend):
    from slmsuite_b200 import CompressedSpotHologram

    g = load(name)
    _basis, method, maxiter, kw = RUN[name]
    h = build(CompressedSpotHologram, g, name)
    assert np.abs(h.spot_zernike - g["spot_zernike"]).max() < 1e-9
    assert np.array_equal(h.zernike_basis, g["zernike_basis"])
    assert rel(h.target, g["target"]) <= 1e-6
    h.optimize(method, maxiter=maxiter, verbose=False, **kw)
    check(h, g)
    assert h._lib.slmgs_comp_launch_count(h._ctx) >= 4 * maxiter


@pytest.mark.parametrize("name", ["compressed_3d_kim", "compressed_mraf_leonardo"])
def test_callback_path_agrees(name, backend):
    from slmsuite_b200 import CompressedSpotHologram

    g = load(name)
    _basis, method, maxiter, kw = RUN[name]
    h = build(CompressedSpotHologram, g, name)
    seen = []

    def cb(holo):
        seen.append(holo.amp_ff.copy())
        return False

    h.optimize(method, maxiter=maxiter, verbose=False, callback=cb, **kw)
    assert len(seen) == maxiter and seen[0].shape == (len(h),)
    check(h, g)


def test_moving_the_spots_between_calls(backend):
    """_spots.py:638-650: spot_zernike may be edited in place; the kernels follow."""
    from slmsuite_b200 import CompressedSpotHologram

    g = load("compressed_2d_leonardo")
    h = build(CompressedSpotHologram, g, "compressed_2d_leonardo")
    h.optimize("GS", maxiter=2, verbose=False)
    a0 = h.amp_ff.copy()
    h.spot_zernike[0] += 5.0
    h.optimize("GS", maxiter=1, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = build(compressed_oracle.OracleCompressedSpotHologram, g, "compressed_2d_leonardo")
        o.optimize("GS", maxiter=2, verbose=False)
        o.spot_zernike[0] += 5.0
        o._kernel = None
        o.optimize("GS", maxiter=1, verbose=False)
    assert rel(h.amp_ff, o.amp_ff) <= 1e-5
    assert rel(h.amp_ff, a0) > 1e-3


def test_errors(backend):
    from slmsuite_b200 import CompressedSpotHologram

    g = load("compressed_2d_leonardo")
    grid = (g["x_grid"], g["y_grid"])
    with pytest.raises(ValueError):
        CompressedSpotHologram(g["spot_vectors"])  # cameraslm must be passed
    with pytest.raises(ValueError):
        CompressedSpotHologram(g["spot_vectors"], basis=[4, 3], slm_grid=grid, zernike_scaling=1.0)  # no x, y
    with pytest.raises(ValueError):
        CompressedSpotHologram(g["spot_vectors"], spot_amp=np.ones(3), slm_grid=grid, zernike_scaling=1.0)
    with pytest.raises(ValueError):
        CompressedSpotHologram(np.zeros((14, 4)), basis=list(range(1, 15)), slm_grid=grid, zernike_scaling=1.0)  # > 10 basis terms
    h = build(CompressedSpotHologram, g, "compressed_2d_leonardo")
    with pytest.raises(NameError):
        h.get_padded_shape((64, 64))
    with pytest.raises(NotImplementedError):
        h.optimize("WGS-Leonardo", maxiter=2, verbose=False, feedback="experimental_spot")
    with pytest.raises(ValueError):
        h.optimize("nope", maxiter=1, verbose=False)


def test_accepts_the_reference_cameraslm_object(emu):
    """Drop-in check where the reference tree is present (build container only): the reference's own FourierSLM object
    is passed as ``cameraslm=`` to both classes and the results agree."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference tree not present")
    ref_loader.load_reference()
    from slmsuite.hardware.cameras.simulated import SimulatedCamera
    from slmsuite.hardware.cameraslms import FourierSLM
    from slmsuite.hardware.slms.simulated import SimulatedSLM
    from slmsuite.holography.algorithms import CompressedSpotHologram as RefCompressed

    from slmsuite_b200 import CompressedSpotHologram

    rng = np.random.default_rng(17)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        slm = SimulatedSLM((72, 56), pitch_um=(8, 8))
        M = np.array([[4000.0, 150.0], [-120.0, 3800.0]])
        b = np.array([[400.0], [300.0]])
        cam = SimulatedCamera(slm, resolution=(800, 600), M=M, b=b, bitdepth=8)
        fs = FourierSLM(cam, slm)
        fs.calibrations["fourier"] = {"M": M, "b": b, "a": np.array([[0.0], [0.0]])}
        v = np.vstack([rng.uniform(-0.02, 0.02, (2, 9)), rng.uniform(-2e-4, 2e-4, (1, 9))])
        phase = rng.uniform(-np.pi, np.pi, slm.shape).astype(np.float32)
        ref = RefCompressed(v.copy(), basis="kxy", cameraslm=fs)
        ref.reset_phase(phase)
        ref.reset(reset_phase=False)
        ref.optimize("WGS-Leonardo", maxiter=5, verbose=False)
        mine = CompressedSpotHologram(v.copy(), basis="kxy", cameraslm=fs, phase=phase)
        mine.optimize("WGS-Leonardo", maxiter=5, verbose=False)
    assert np.abs(mine.spot_zernike - ref.spot_zernike).max() < 1e-9
    assert np.array_equal(mine.zernike_basis, ref.zernike_basis)
    assert tuple(mine.shape) == tuple(int(s) for s in ref.shape)
    assert rel(mine.amp_ff, ref.amp_ff) <= 1e-5
    assert rel(mine.weights, ref.weights) <= 1e-5
    assert phase_rms(mine.phase, ref.phase) <= 1e-4
