"""
CPU oracle for the simulated-camera image  --  TEST INFRASTRUCTURE ONLY.

NumPy / SciPy restatement of ``SimulatedCamera._get_image_hw`` (slmsuite/hardware/cameras/simulated.py:344-402,
noise=None): the far field of the SLM's displayed phase, ``|.|^2``, nearest-neighbour resampling onto the camera
pixels (``scipy.ndimage.map_coordinates(order=0)``) or a centred crop, exposure * gain, clipping, integer cast.
SURVEY.md 8f rank 3.

Parity status: PINNED.  ``oracle/make_golden_camera.py`` builds the unmodified reference's ``SimulatedSLM`` +
``SimulatedCamera``, records the images it returns (tests/golden/camera_*.npz, together with the geometry the
reference derived) and checks this restatement against them bit for bit; tests/test_camera.py replays them.
"""
import numpy as np
from scipy.ndimage import map_coordinates

from oracle import gs_oracle


def phase_from_display(display, slm_bitresolution, phase_sim, dtype=np.float32):
    """simulated.py:365-366."""
    phase = -np.asarray(display).astype(dtype) * (2 * np.pi / slm_bitresolution)
    return phase - phase.min() + np.asarray(phase_sim).astype(dtype)


def camera_image(phase, amp, slm_shape, shape_padded, cam_shape, knm_cam, exposure_s, gain, bitdepth):
    """
    ``phase`` float32 (already prepared, see phase_from_display); ``amp`` = source["amplitude_sim"], used RAW
    (simulated.py:364 overwrites Hologram.amp without normalising); ``knm_cam`` None = centred crop (:377-379).
    Returns (integer image, float32 image before clipping).
    """
    h = gs_oracle.OracleHologram(tuple(int(s) for s in shape_padded), amp=np.array(amp, dtype=np.float32),
                                 phase=np.zeros(slm_shape, np.float32), slm_shape=tuple(slm_shape))
    h.amp = np.array(amp, dtype=np.float32)
    h.reset_phase(np.asarray(phase, dtype=np.float32))
    h._forward()                       # farfield = fftshift(fft2(fftshift(nearfield), norm="ortho")), :1048
    ff = h.farfield
    if knm_cam is not None:
        img = map_coordinates(np.abs(ff) ** 2, knm_cam, order=0)
    else:
        img = np.abs(ff) ** 2
        i0, i1, i2, i3 = gs_oracle.crop_bounds(ff.shape, cam_shape)
        img = img[i0:i1, i2:i3]
    img = img * np.float32(1.0)
    img *= exposure_s * gain
    raw = img.copy()
    bitresolution = 2 ** int(bitdepth)
    img[img > bitresolution - 1] = bitresolution - 1
    return img.astype(np.uint8 if bitdepth <= 8 else np.uint16), raw
