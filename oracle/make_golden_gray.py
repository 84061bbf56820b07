"""
Generate tests/golden/gray_*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_gray.py

Runs ``SimulatedSLM.set_phase`` (slmsuite/hardware/slms/slm.py:438-690) on seeded ``get_phase()``-like
inputs (float32 phase + pi) for several bit depths, with and without a wavefront correction, stores the
``display`` arrays, and checks ``oracle.gray_oracle.phase2gray`` against them.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gray_oracle, ref_loader  # noqa: E402

CASES = {
    # name: (shape (h, w), bitdepth, seed, with_correction)
    "gray_8bit_96x160": ((96, 160), 8, 1, False),
    "gray_10bit_64x64": ((64, 64), 10, 2, False),
    "gray_12bit_corr_48x80": ((48, 80), 12, 3, True),
    "gray_8bit_corr_37x61": ((37, 61), 8, 4, True),
}


def inputs(shape, seed, corr):
    rng = np.random.default_rng(seed)
    raw = rng.uniform(-np.pi, np.pi, shape).astype(np.float32)
    raw.flat[0] = -np.pi          # edge: get_phase() == 0
    raw.flat[1] = np.float32(np.pi)
    correction = rng.uniform(-8, 8, shape) if corr else None   # float64, like source["phase"]
    return raw, correction


def main():
    ref_loader.load_reference()
    from slmsuite.hardware.slms.simulated import SimulatedSLM

    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (shape, bitdepth, seed, corr) in CASES.items():
        raw, correction = inputs(shape, seed, corr)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            slm = SimulatedSLM((shape[1], shape[0]), bitdepth=bitdepth)
            if corr:
                slm.source["phase"] = correction
            display = np.array(slm.set_phase(raw + np.pi, phase_correct=corr, settle=False))
        mine = gray_oracle.phase2gray(raw + np.pi, bitdepth, correction)
        ok = np.array_equal(mine, display) and mine.dtype == display.dtype
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), display=display, bitdepth=np.int64(bitdepth))
        print(f"{name:28s} dtype {display.dtype} oracle == reference: {ok}")
        assert ok


if __name__ == "__main__":
    main()
