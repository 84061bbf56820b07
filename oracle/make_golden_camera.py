"""
Generate tests/golden/camera_*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_camera.py

Builds the reference's ``SimulatedSLM`` + ``SimulatedCamera`` (hardware/slms/simulated.py,
hardware/cameras/simulated.py), displays a seeded phase, records ``cam.get_image()`` together with the inputs the
B200 path needs (SLM gray levels, source amplitude / phase, affine map, the reference's own ``knm_cam`` and padded
shape), and checks ``oracle.camera_oracle`` against the recorded image.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import camera_oracle, ref_loader  # noqa: E402

CASES = {
    # name: slm (w, h), slm bitdepth, cam resolution (w, h) or None, cam bitdepth, M, b, exposure, gain, seed, source
    "camera_affine_8bit": dict(slm=(96, 64), slm_bits=8, res=(140, 100), cam_bits=8,
                               M=[[900.0, 40.0], [-30.0, 850.0]], b=[[70.0], [50.0]], exposure=3000.0, gain=1, seed=0,
                               source=False),
    "camera_affine_12bit_source": dict(slm=(128, 80), slm_bits=10, res=(90, 120), cam_bits=12,
                                       M=[[500.0, -120.0], [100.0, 620.0]], b=[[40.0], [65.0]], exposure=40000.0, gain=2,
                                       seed=1, source=True),
    "camera_beyond_kspace_8bit": dict(slm=(64, 64), slm_bits=8, res=(80, 80), cam_bits=8,
                                      M=[[600.0, 0.0], [0.0, 600.0]], b=[[40.0], [40.0]], exposure=2000.0, gain=1, seed=2,
                                      source=False),
    "camera_crop_12bit": dict(slm=(128, 64), slm_bits=8, res=None, cam_bits=12, M=None, b=None, exposure=30000.0,
                              gain=1, seed=3, source=True),
}


def build_reference(case):
    from slmsuite.hardware.cameras.simulated import SimulatedCamera
    from slmsuite.hardware.slms.simulated import SimulatedSLM

    rng = np.random.default_rng(case["seed"])
    w, h = case["slm"]
    kw = {}
    if case["source"]:
        yy, xx = np.mgrid[-1:1:h * 1j, -1:1:w * 1j]
        kw["source"] = {"amplitude_sim": np.exp(-(xx ** 2 + yy ** 2) / 0.7), "phase_sim": 0.8 * (xx ** 2 - yy ** 2)}
    slm = SimulatedSLM((w, h), pitch_um=(8, 8), bitdepth=case["slm_bits"], **kw)
    M = None if case["M"] is None else np.array(case["M"])
    b = None if case["b"] is None else np.array(case["b"])
    cam = SimulatedCamera(slm, resolution=case["res"], M=M, b=b, bitdepth=case["cam_bits"], gain=case["gain"])
    # a blazed grating plus noise: a bright first order somewhere on the camera and a speckle background
    yy, xx = np.mgrid[0:h, 0:w]
    phase = 2 * np.pi * (0.11 * xx - 0.07 * yy) + rng.uniform(0, 1.5, (h, w))
    slm.set_phase(phase, settle=False, phase_correct=False)
    cam.set_exposure(case["exposure"])
    return slm, cam


def main():
    ref_loader.load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, case in CASES.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            slm, cam = build_reference(case)
            img = np.array(cam.get_image())
        display = np.array(slm.display)
        amp = np.array(slm.source["amplitude_sim"], dtype=np.float64)
        phase_sim = np.array(slm.source["phase_sim"], dtype=np.float64)
        interp = bool(cam._interpolate)
        knm = np.array(cam.knm_cam) if interp else np.zeros((0,))
        phase = camera_oracle.phase_from_display(display, slm.bitresolution, phase_sim)
        mine, raw = camera_oracle.camera_image(phase, amp, slm.shape, cam.shape_padded, cam.shape,
                                               knm if interp else None, cam.exposure_s, cam.gain, case["cam_bits"])
        ok = np.array_equal(mine, img) and mine.dtype == img.dtype
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"), image=img, display=display, amp=amp, phase_sim=phase_sim,
            knm_cam=knm, shape_padded=np.array(cam.shape_padded, dtype=np.int64), interpolate=np.int64(interp),
            slm_pitch=np.array(slm.pitch, dtype=np.float64), slm_bitresolution=np.int64(slm.bitresolution),
            cam_bitdepth=np.int64(case["cam_bits"]), exposure=np.float64(cam.exposure_s), gain=np.float64(cam.gain),
            M=np.zeros((0,)) if case["M"] is None else np.array(case["M"], dtype=np.float64),
            b=np.zeros((0,)) if case["b"] is None else np.array(case["b"], dtype=np.float64),
            resolution=np.array(cam.shape[::-1], dtype=np.int64))
        frac = float((img > 0).mean())
        sat = float((img == 2 ** case["cam_bits"] - 1).mean())
        print(f"{name:30s} {img.shape} {img.dtype} padded {tuple(int(s) for s in cam.shape_padded)} interp {interp} "
              f"nonzero {frac:.2f} saturated {sat:.3f} oracle == reference: {ok}")
        assert ok


if __name__ == "__main__":
    main()
