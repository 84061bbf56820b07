"""
CPU oracle for the phase -> SLM gray-level conversion  --  TEST INFRASTRUCTURE ONLY.

NumPy restatement of ``SLM.set_phase`` / ``SLM._phase2gray`` for ``phase_scaling == 1``
(slmsuite/hardware/slms/slm.py:636-690 and :695-743), the step that follows
``Hologram.get_phase()`` in every real use (hardware/cameraslms.py:1153).  SURVEY.md 8f rank 2.

Parity status: PINNED.  ``oracle/make_golden_gray.py`` runs the unmodified reference
(``SimulatedSLM.set_phase``) and stores its ``display`` arrays under tests/golden/gray_*.npz;
tests/test_gray.py replays them bit for bit.
"""

import numpy as np


def phase2gray(phase, bitdepth, phase_correction=None):
    """
    ``display`` for a float phase array (radians), as ``SLM.set_phase(phase)`` computes it.

    slm.py:211-213  dtype uint8 for bitdepth <= 8 else uint16
    slm.py:217,676  the phase is copied into a float64 cache; source["phase"] is added when phase_correct
    slm.py:725-743  scale by -(bitresolution / 2 pi), shift negative, rint, unsafe cast, -1, bit mask
    """
    bitresolution = 2 ** int(bitdepth)
    dtype = np.uint8 if bitdepth <= 8 else np.uint16
    ph = np.zeros(np.shape(phase))  # float64 cache
    np.copyto(ph, phase)
    if phase_correction is not None:
        ph += np.asarray(phase_correction)
    factor = -(bitresolution / 2 / np.pi)
    ph *= factor
    maximum = np.amax(ph)
    if maximum >= 0:
        ph -= bitresolution * 2 * float(np.ceil(maximum / bitresolution))
    np.rint(ph, out=ph)
    out = np.zeros(ph.shape, dtype=dtype)
    with np.errstate(invalid="ignore"):
        np.copyto(out, ph, casting="unsafe")
    out -= 1
    np.bitwise_and(out, int(bitresolution - 1), out=out)
    return out
