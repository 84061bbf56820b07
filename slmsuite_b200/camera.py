"""
What a simulated camera sees of a hologram: far field -> |.|^2 -> nearest-neighbour resampling onto camera pixels
-> exposure / gain -> clipping -> integer readout.  "Next" row 3 of SURVEY.md 8f: the compute part of
``slmsuite.hardware.cameras.simulated.SimulatedCamera`` (hardware/cameras/simulated.py:72-402), i.e. the step
AROUND the GS/WGS path in closed-loop simulation (``experimental_spot`` feedback and the Fourier calibration in
the reference's tests look at the SLM through it).

The reference class needs a ``SimulatedSLM`` hardware object; this mirror takes the few numbers it reads from it
(shape, normalised pitch, bit depth, source amplitude / phase) so that it stays inside the hot-path scope:

    cam = SimulatedCamera(slm_shape, resolution, M=M, b=b, slm_pitch=(dx, dy), bitdepth=8)
    cam.set_exposure(3000.0)
    img = cam.get_image(display, slm_bitresolution=256)        # display = SLM gray levels (uint8/uint16)
    img = cam.get_image_from_phase(phase)                      # or an analog phase

Everything after the host-side phase preparation runs on the device through the C ABI
(``slmgs_set_sample_grid`` / ``slmgs_sample_intensity``); there is no CPU fallback.
"""
import ctypes as C
import warnings

import numpy as np

from . import _lib
from .hologram import Hologram


def padded_shape_for_precision(slm_shape, slm_pitch, precision, padding_order=1, square_padding=True):
    """``Hologram.get_padded_shape(slm, precision=..., precision_basis="kxy")``, _hologram.py:686-725, with the SLM
    object replaced by its shape and normalised pitch (``slm.pitch`` = pitch / wavelength)."""
    if precision <= 0:
        raise ValueError("Precision passed to get_padded_shape() must be positive.")
    fs = 1 / np.amin(slm_pitch)  # sampling frequency
    pixels = fs / precision
    pixels = np.power(2, int(np.ceil(np.log2(pixels))))
    precision_shape = (pixels, pixels)
    if padding_order > 0:
        padding_shape = np.power(2, np.ceil(np.log2(slm_shape)) + padding_order - 1).astype(int)
    else:
        padding_shape = slm_shape
    shape = tuple(np.amax(np.vstack((precision_shape, padding_shape)), axis=0))
    if square_padding:
        largest = np.amax(shape)
        shape = (largest, largest)
    return tuple(int(s) for s in shape)


def camera_knm_grid(resolution, M, b, slm_shape, slm_pitch):
    """
    ``SimulatedCamera.set_affine`` (simulated.py:156-186): the camera's pixel grid transformed to the SLM's k-space
    (``toolbox.transform_grid(..., direction="rev")``, toolbox/__init__.py:1577-1587), the padded shape that resolves
    one camera pixel, and the camera pixel centres in far-field pixel coordinates ``knm_cam`` (2, rows, cols; (y, x)
    order).  ``resolution`` is (width, height) like the reference's.
    """
    gx, gy = np.meshgrid(np.arange(resolution[0]), np.arange(resolution[1]))
    M = np.squeeze(np.asarray(M, dtype=float))
    if M.shape != (2, 2):
        raise ValueError("Expected transform to be None, scalar, or a 2x2 matrix.")
    shift = np.squeeze(np.asarray(b, dtype=float))
    inv = np.linalg.inv(M)
    kx = inv[0, 0] * (gx - shift[0]) + inv[0, 1] * (gy - shift[1])
    ky = inv[1, 0] * (gx - shift[0]) + inv[1, 1] * (gy - shift[1])
    # Fourier space must be sufficiently padded to resolve the camera pixels (simulated.py:168-175)
    dkxy = np.sqrt((kx[:2, :2] - kx[0, 0]) ** 2 + (ky[:2, :2] - ky[0, 0]) ** 2)
    dkxy_min = dkxy.ravel()[1:].min()
    shape_padded = padded_shape_for_precision(slm_shape, slm_pitch, dkxy_min)
    # kxy -> knm; the reference scales both axes with pitch[1] (simulated.py:180-185), reproduced as is
    knm_cam = np.array([
        shape_padded[0] * slm_pitch[1] * ky + shape_padded[0] / 2,
        shape_padded[1] * slm_pitch[1] * kx + shape_padded[1] / 2,
    ])
    return shape_padded, knm_cam


class SimulatedCamera:
    """
    Device-side mirror of ``SimulatedCamera._get_image_hw`` (hardware/cameras/simulated.py:344-402).

    Parameters
    ----------
    slm_shape : (int, int)
        ``slm.shape`` (rows, columns).
    resolution : (int, int) or None
        Camera (width, height); ``None`` = the SLM's, which makes the camera a centred crop of the un-padded far
        field (simulated.py:113-116, :377-379) -- the SLM shape must then be a power of two per axis here.
    M, b : array_like or None
        Affine map k-space -> camera pixels (simulated.py:130-186).  Both ``None``: no interpolation.
    slm_pitch : (float, float)
        Normalised SLM pitch (pitch / wavelength), ``slm.pitch``.
    bitdepth : int
        Camera bit depth: ``bitresolution = 2**bitdepth``, dtype uint8 / uint16 (cameras/camera.py).
    amp, phase_sim : array_like or None
        ``slm.source["amplitude_sim"]`` / ``["phase_sim"]``; default ones / zeros like ``SimulatedSLM``'s.
    noise : dict or None
        ``{"dark": fn, "read": fn}`` applied on the host exactly as the reference does (:384-396).
    knm_cam, shape_padded : optional
        Precomputed geometry (e.g. taken from a reference ``SimulatedCamera``) instead of ``M``, ``b``, ``slm_pitch``.
    """

    def __init__(self, slm_shape, resolution=None, M=None, b=None, slm_pitch=None, bitdepth=8, amp=None,
                 phase_sim=None, noise=None, gain=1, knm_cam=None, shape_padded=None, device=0):
        self.slm_shape = tuple(int(s) for s in slm_shape)
        if resolution is None:
            resolution = self.slm_shape[::-1]
        self.shape = (int(resolution[1]), int(resolution[0]))
        self.bitdepth = int(bitdepth)
        self.bitresolution = 2 ** self.bitdepth
        if self.bitdepth > 16:
            raise ValueError("bitdepth > 16 is not supported")
        self.dtype = np.uint8 if self.bitdepth <= 8 else np.uint16
        self.gain = gain
        self.noise = noise
        self.exposure_s = 1.0
        self.amp = np.ones(self.slm_shape) if amp is None else np.asarray(amp)
        self.phase_sim = np.zeros(self.slm_shape) if phase_sim is None else np.asarray(phase_sim)

        if knm_cam is not None:
            if shape_padded is None:
                raise ValueError("shape_padded must accompany knm_cam")
            self._interpolate = True
            self.shape_padded = tuple(int(s) for s in shape_padded)
            self.knm_cam = np.asarray(knm_cam, dtype=np.float64)
        elif M is None or b is None:
            # aligned with the SLM grid: the image is unpad(|farfield|^2, cam.shape) of the un-padded transform
            self._interpolate = False
            self.shape_padded = self.slm_shape
            if self.shape[0] > self.slm_shape[0] or self.shape[1] > self.slm_shape[1]:
                raise ValueError("camera without an affine transform cannot be larger than the SLM")
            i0 = (self.slm_shape[0] - self.shape[0]) // 2  # toolbox.unpad, toolbox/__init__.py:1701-1709
            j0 = (self.slm_shape[1] - self.shape[1]) // 2
            yy, xx = np.meshgrid(np.arange(self.shape[0]) + i0, np.arange(self.shape[1]) + j0, indexing="ij")
            self.knm_cam = np.array([yy, xx], dtype=np.float64)
        else:
            if slm_pitch is None:
                raise ValueError("slm_pitch (normalised SLM pixel pitch) is needed to place the camera in k-space")
            self._interpolate = True
            self.shape_padded, self.knm_cam = camera_knm_grid(resolution, M, b, self.slm_shape, slm_pitch)
            if (np.amax(np.abs(self.knm_cam[0] - self.shape_padded[0] / 2)) > self.shape_padded[1] / 2 or
                    np.amax(np.abs(self.knm_cam[1] - self.shape_padded[1] / 2)) > self.shape_padded[0] / 2):
                warnings.warn("Camera extends beyond the accessible SLM k-space; some pixels may not be targetable.")
        if self.knm_cam.shape != (2,) + self.shape:
            raise ValueError(f"knm_cam of shape {self.knm_cam.shape} does not match the camera shape {self.shape}")

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=Warning)
            self._hologram = Hologram(self.shape_padded, amp=self.amp, phase=np.zeros(self.slm_shape, np.float32),
                                      slm_shape=self.slm_shape, device=device)
        h = self._hologram
        ky = np.ascontiguousarray(self.knm_cam[0].ravel(), dtype=np.float64)
        kx = np.ascontiguousarray(self.knm_cam[1].ravel(), dtype=np.float64)
        h._check(h._lib.slmgs_set_sample_grid(h._ctx, ky.size, _lib.dptr(ky), _lib.dptr(kx)))

    # exposure, simulated.py:336-342
    def get_exposure(self):
        return self.exposure_s

    def set_exposure(self, exposure_s):
        self.exposure_s = exposure_s

    def _image(self, phase):
        h = self._hologram
        # simulated.py:364-366: amp is overwritten raw (not normalised), then reset_phase
        h.amp = np.array(self.amp, dtype=h.dtype)
        h.reset_phase(phase)
        scale = np.float32(self.exposure_s * self.gain)
        if self.noise is None:
            out = np.empty(self.shape, dtype=self.dtype)
            kind = 1 if self.dtype == np.uint8 else 2
            h._check(h._lib.slmgs_sample_intensity(h._ctx, scale, np.float32(self.bitresolution - 1), kind,
                                                   out.ctypes.data_as(C.c_void_p)))
            return out
        img = np.empty(self.shape, dtype=np.float32)
        h._check(h._lib.slmgs_sample_intensity(h._ctx, scale, np.float32(-1.0), 0, img.ctypes.data_as(C.c_void_p)))
        for key in self.noise.keys():  # simulated.py:384-396
            if key == "dark":
                img = img + self.noise["dark"](np.ones_like(img) * self.bitresolution) / self.exposure_s
            elif key == "read":
                img = img + self.noise["read"](np.ones_like(img) * self.bitresolution)
            else:
                raise RuntimeError("Unknown noise source %s specified!" % key)
        img[img > self.bitresolution - 1] = self.bitresolution - 1
        return img.astype(self.dtype)

    def get_image(self, display, slm_bitresolution):
        """Image for the SLM gray levels ``display`` (the quantised phase, simulated.py:365-366)."""
        dt = self._hologram.dtype
        phase = -np.asarray(display).astype(dt) * (2 * np.pi / slm_bitresolution)
        return self._image(phase - phase.min() + self.phase_sim.astype(dt))

    def get_image_from_phase(self, phase):
        """Image for an analog phase (the commented alternative at simulated.py:361)."""
        dt = self._hologram.dtype
        return self._image(np.asarray(phase, dtype=dt) + self.phase_sim.astype(dt))

    def get_farfield_intensity(self):
        """float32 |farfield|^2 at the camera pixels of the last image's phase, no exposure / clipping."""
        h = self._hologram
        img = np.empty(self.shape, dtype=np.float32)
        h._check(h._lib.slmgs_sample_intensity(h._ctx, np.float32(1.0), np.float32(-1.0), 0,
                                               img.ctypes.data_as(C.c_void_p)))
        return img
