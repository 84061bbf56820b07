"""
Generate tests/golden/compressed_*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_compressed.py

Runs the reference's ``CompressedSpotHologram`` (NumPy backend, _spots.py:178-1019) on a ``SimulatedSLM`` +
``SimulatedCamera`` + ``FourierSLM`` with an explicit Fourier calibration, from an explicit seeded phase, records
inputs (spot vectors, SLM grid, aperture scaling, amplitude, phase) and results, and checks
``oracle.compressed_oracle`` against them (rel-RMSE(|farfield|) <= 1e-5, phase rms <= 1e-4 rad).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import compressed_oracle, ref_loader  # noqa: E402

CASES = {
    "compressed_2d_leonardo": dict(slm=(96, 64), n=12, dim=2, basis="kxy", method="WGS-Leonardo", maxiter=6, seed=0, kw={}),
    "compressed_3d_kim": dict(slm=(80, 48), n=9, dim=3, basis="kxy", method="WGS-Kim", maxiter=8, seed=1,
                              kw={"fix_phase_iteration": 3}),
    "compressed_zernike5_gs": dict(slm=(64, 64), n=7, dim=5, basis=[2, 1, 4, 3, 5], method="GS", maxiter=5, seed=2, kw={}),
    "compressed_mraf_leonardo": dict(slm=(96, 64), n=10, dim=2, basis="kxy", method="WGS-Leonardo", maxiter=6, seed=3,
                                     kw={}, mraf=True),
    "compressed_2d_nogrette": dict(slm=(64, 96), n=8, dim=2, basis="kxy", method="WGS-Nogrette", maxiter=5, seed=4, kw={}),
}


def inputs(case):
    rng = np.random.default_rng(case["seed"])
    n, dim = case["n"], case["dim"]
    if case["basis"] == "kxy":
        v = rng.uniform(-0.02, 0.02, (dim, n))
        if dim == 3:
            v[2] = rng.uniform(-2e-4, 2e-4, n)   # focal power
    else:
        v = rng.uniform(-20, 20, (dim, n))
        v[2:] = rng.uniform(-3, 3, (dim - 2, n))
    amp = rng.uniform(0.5, 1.5, n)
    if case.get("mraf"):
        amp[2] = np.nan
        amp[5] = 0.0
        amp[7] = np.nan
    w, h = case["slm"]
    phase = rng.uniform(-np.pi, np.pi, (h, w)).astype(np.float32)
    return v, amp, phase


def main():
    ref_loader.load_reference()
    from slmsuite.hardware.cameras.simulated import SimulatedCamera
    from slmsuite.hardware.cameraslms import FourierSLM
    from slmsuite.hardware.slms.simulated import SimulatedSLM
    from slmsuite.holography.algorithms import CompressedSpotHologram

    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, case in CASES.items():
        v, spot_amp, phase0 = inputs(case)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            slm = SimulatedSLM(case["slm"], pitch_um=(8, 8))
            M = np.array([[4000.0, 150.0], [-120.0, 3800.0]])
            b = np.array([[400.0], [300.0]])
            cam = SimulatedCamera(slm, resolution=(800, 600), M=M, b=b, bitdepth=8)
            fs = FourierSLM(cam, slm)
            fs.calibrations["fourier"] = {"M": M, "b": b, "a": np.array([[0.0], [0.0]])}
            h = CompressedSpotHologram(v, basis=case["basis"], spot_amp=spot_amp.copy(), cameraslm=fs)
            h.reset_phase(phase0)
            h.reset(reset_phase=False)
            h.optimize(case["method"], maxiter=case["maxiter"], verbose=False, **case["kw"])
        scaling = float(slm.get_source_zernike_scaling())
        grid = (np.array(slm.grid[0]), np.array(slm.grid[1]))
        amp = np.array(slm._get_source_amplitude(), dtype=np.float64)
        gold = dict(
            spot_vectors=v, spot_amp=spot_amp, phase0=phase0, x_grid=grid[0], y_grid=grid[1],
            zernike_scaling=np.float64(scaling), amp=amp, spot_zernike=np.array(h.spot_zernike),
            zernike_basis=np.array(h.zernike_basis, dtype=np.int64),
            phase=np.array(h.phase), farfield=np.array(h.farfield), amp_ff=np.array(h.amp_ff),
            weights=np.array(h.weights), target=np.array(h.target), iter=np.int64(h.iter),
            fixed_phase=np.int64(bool(h.flags.get("fixed_phase", False))),
        )
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = compressed_oracle.OracleCompressedSpotHologram(
                v, basis=case["basis"], spot_amp=spot_amp.copy(), slm_grid=grid, zernike_scaling=scaling, amp=amp,
                phase=phase0)
            o.optimize(case["method"], maxiter=case["maxiter"], verbose=False, **case["kw"])
        ez = np.abs(o.spot_zernike - gold["spot_zernike"]).max()
        ea = np.linalg.norm(o.amp_ff - gold["amp_ff"]) / np.linalg.norm(gold["amp_ff"])
        dphi = np.angle(np.exp(1j * (o.phase.astype(np.float64) - gold["phase"])))
        ep = float(np.sqrt(np.mean(dphi ** 2)))
        m = ~np.isnan(gold["weights"])
        ew = np.linalg.norm((o.weights - gold["weights"])[m]) / np.linalg.norm(gold["weights"][m])
        print(f"{name:28s} N={len(spot_amp):3d} basis {gold['zernike_basis'].tolist()} iter {int(h.iter)} "
              f"zernike coeff diff {ez:.1e} amp_ff rel-rmse {ea:.1e} phase rms {ep:.1e} weights {ew:.1e} "
              f"fixed {bool(o.flags.get('fixed_phase'))}/{bool(gold['fixed_phase'])}")
        assert ez < 1e-9 and ea <= 1e-5 and ep <= 1e-4 and ew <= 1e-5
        assert int(o.iter) == int(h.iter) and bool(o.flags.get("fixed_phase", False)) == bool(gold["fixed_phase"])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **gold)


if __name__ == "__main__":
    main()
