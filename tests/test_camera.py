"""
Simulated-camera image ("next" row 3): ``slmsuite_b200.camera.SimulatedCamera`` against the images
the UNMODIFIED reference's ``SimulatedCamera.get_image()`` returned (tests/golden/camera_*.npz, made by
oracle/make_golden_camera.py from hardware/cameras/simulated.py:344-402).

Tolerances.  The far field is float32 on both sides (rel-RMSE <= 1e-5 on the un-clipped float image); the final
image is an integer TRUNCATION of that float image, so a pixel may differ by one gray level only where the
reference's float value sits within 1e-3 (relative) of an integer; everywhere else it must be identical.
"""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import camera_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "camera_*.npz")))


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def oracle_image(g):
    interp = bool(g["interpolate"])
    phase = camera_oracle.phase_from_display(g["display"], int(g["slm_bitresolution"]), g["phase_sim"])
    return camera_oracle.camera_image(phase, g["amp"], g["display"].shape, g["shape_padded"], g["image"].shape,
                                      g["knm_cam"] if interp else None, float(g["exposure"]), float(g["gain"]),
                                      int(g["cam_bitdepth"]))


def test_golden_cases_exist():
    assert len(NAMES) >= 4


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_images(name):
    g = load(name)
    img, _raw = oracle_image(g)
    assert img.dtype == g["image"].dtype
    assert np.array_equal(img, g["image"])


@pytest.mark.parametrize("name", NAMES)
def test_geometry_restatement_matches_reference(name):
    """camera_knm_grid / padded_shape_for_precision against the reference's own knm_cam and shape_padded."""
    from slmsuite_b200.camera import camera_knm_grid

    g = load(name)
    if not bool(g["interpolate"]):
        pytest.skip("no affine map in this case")
    shape_padded, knm = camera_knm_grid(tuple(g["resolution"]), g["M"], g["b"], g["display"].shape, g["slm_pitch"])
    assert tuple(shape_padded) == tuple(int(s) for s in g["shape_padded"])
    assert np.array_equal(knm, g["knm_cam"])


def make_camera(g, from_affine):
    from slmsuite_b200.camera import SimulatedCamera

    interp = bool(g["interpolate"])
    kw = dict(bitdepth=int(g["cam_bitdepth"]), amp=g["amp"], phase_sim=g["phase_sim"], gain=float(g["gain"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if not interp:
            cam = SimulatedCamera(g["display"].shape, **kw)
        elif from_affine:
            cam = SimulatedCamera(g["display"].shape, resolution=tuple(g["resolution"]), M=g["M"], b=g["b"],
                                  slm_pitch=g["slm_pitch"], **kw)
        else:
            cam = SimulatedCamera(g["display"].shape, resolution=tuple(g["resolution"]), knm_cam=g["knm_cam"],
                                  shape_padded=g["shape_padded"], **kw)
    cam.set_exposure(float(g["exposure"]))
    return cam


@pytest.mark.parametrize("from_affine", [True, False])
@pytest.mark.parametrize("name", NAMES)
def test_camera_image_matches_reference(name, from_affine, backend):
    g = load(name)
    cam = make_camera(g, from_affine)
    img = cam.get_image(g["display"], int(g["slm_bitresolution"]))
    gold = g["image"]
    assert img.dtype == gold.dtype and img.shape == gold.shape
    _gold_img, raw = oracle_image(g)
    # float image before clipping
    mine = cam.get_farfield_intensity() * np.float32(float(g["exposure"]) * float(g["gain"]))
    err = np.linalg.norm(mine.astype(np.float64) - raw) / np.linalg.norm(raw)
    assert err <= 1e-5, err
    assert np.array_equal(mine == 0, raw == 0)  # the same pixels fall outside the SLM's k-space
    # integer image: identical except at float values that sit on a truncation boundary
    diff = img.astype(np.int64) - gold.astype(np.int64)
    near_boundary = np.abs(raw - np.rint(raw)) <= 1e-3 * np.maximum(raw, 1.0)
    assert np.all(np.abs(diff) <= 1)
    assert not np.any((diff != 0) & ~near_boundary)
    assert (diff != 0).mean() < 1e-2


def test_noise_path_runs_on_host_like_the_reference(backend):
    g = load("camera_affine_8bit")
    cam = make_camera(g, True)
    cam.noise = {"read": lambda img: 0.0 * img + 3.0}
    noisy = cam.get_image(g["display"], int(g["slm_bitresolution"]))
    cam.noise = None
    clean = cam.get_image(g["display"], int(g["slm_bitresolution"]))
    expect = np.minimum(clean.astype(np.int64) + 3, 255)
    # clean is truncated before the offset in this comparison, the noisy path after it: equal up to one level
    assert np.all(np.abs(noisy.astype(np.int64) - expect) <= 1)
    cam.noise = {"bogus": lambda img: img}
    with pytest.raises(RuntimeError):
        cam.get_image(g["display"], int(g["slm_bitresolution"]))


def test_errors(backend):
    from slmsuite_b200.camera import SimulatedCamera

    with pytest.raises(ValueError):
        SimulatedCamera((64, 64), resolution=(80, 80), M=np.eye(2), b=np.zeros(2))  # no slm_pitch
    with pytest.raises(ValueError):
        SimulatedCamera((64, 96))  # un-padded far field of a non-power-of-two SLM
    with pytest.raises(ValueError):
        SimulatedCamera((64, 64), knm_cam=np.zeros((2, 64, 64)))  # no shape_padded
