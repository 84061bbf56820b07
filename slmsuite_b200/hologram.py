"""
``Hologram``: host-side mirror of ``slmsuite.holography.algorithms.Hologram`` for the GS / WGS
path, backed by the sm_100a library through the C ABI (``include/slmgs.h``).

Same constructor, ``optimize()`` / ``get_phase()`` signatures, flags, statistics layout and error
behaviour as the reference (file:line citations are relative to the reference tree,
slmsuite v0.4.1 @ 39243f08); the arithmetic runs on the GPU.  Differences a user can see:

* device arrays are exposed as properties that download on read and upload on assignment
  (the reference hands out the backend arrays themselves, _hologram.py:70-77);
* only ``dtype`` float32 / complex64 and power-of-two ``shape`` in [16, 8192] are supported
  (the reference merely warns for other shapes, _hologram.py:378-384); anything else raises;
* ``"CG"`` (torch autograd, _hologram.py:1664-1783) is outside this path and raises ``ValueError``.
"""

import ctypes as C
import warnings

import numpy as np

from . import _lib

# slmsuite/holography/algorithms/_header.py:53-81
ALGORITHM_DEFAULTS = {
    "GS": {"feedback": "computational"},
    "WGS-Leonardo": {"feedback": "computational", "feedback_exponent": 0.8},
    "WGS-Kim": {
        "feedback": "computational",
        "fix_phase_efficiency": None,
        "fix_phase_iteration": 10,
        "feedback_exponent": 0.8,
    },
    "WGS-Nogrette": {"feedback": "computational", "feedback_factor": 0.1},
    "WGS-Wu": {"feedback": "computational", "feedback_exponent": 0.5},
    "WGS-tanh": {"feedback": "computational", "feedback_factor": 0.2, "feedback_exponent": 0.5},
}
ALGORITHM_INDEX = {key: i for i, key in enumerate(ALGORITHM_DEFAULTS.keys())}
FEEDBACK_OPTIONS = [
    "computational",
    "computational_spot",
    "experimental",
    "experimental_spot",
    "external_spot",
]


def _norm(a):
    """sqrt(nansum(a^2)) in a's dtype: Hologram._norm, _hologram.py:1980-2011."""
    return np.sqrt(np.nansum(np.square(a)))


def unpad_bounds(shape, slm_shape):
    """Centred crop indices (i0, i1, i2, i3): toolbox.unpad, toolbox/__init__.py:1665-1712."""
    dy = (shape[0] - slm_shape[0]) / 2.0
    dx = (shape[1] - slm_shape[1]) / 2.0
    if dy < 0 or dx < 0:
        raise ValueError(f"Shape {tuple(shape)} is too small to unpad to shape {tuple(slm_shape)}")
    return (int(np.floor(dy)), int(shape[0] - np.ceil(dy)), int(np.floor(dx)), int(shape[1] - np.ceil(dx)))


def calculate_stats(feedback_amp, target_amp, total=None):
    """
    Host version of ``_HologramStats._calculate_stats`` (_stats.py:7-116, mp=np,
    efficiency_compensation=False, raw=False) for the N-vectors of spot statistics.
    Unlike the reference it does not normalise its arguments in place.
    """
    feedback_amp = np.array(feedback_amp, copy=True)
    target_amp = np.array(target_amp, copy=True)
    feedback_pwr = np.square(feedback_amp)
    target_pwr = np.square(target_amp)
    if total is not None:
        efficiency = float(np.nansum(feedback_pwr)) / total
    feedback_norm = np.sum(feedback_pwr)
    feedback_pwr = feedback_pwr * (1 / feedback_norm)
    feedback_amp = feedback_amp * (1 / np.sqrt(feedback_norm))
    target_norm = np.nansum(target_pwr)
    target_pwr = target_pwr * (1 / target_norm)
    target_amp = target_amp * (1 / np.sqrt(target_norm))
    if total is None:
        efficiency = np.square(float(np.nansum(np.multiply(target_amp, feedback_amp))))
    mask = np.logical_and(target_pwr != 0, np.logical_not(np.isnan(target_pwr)))
    fm = feedback_pwr[mask]
    tm = target_pwr[mask]
    ratio = np.divide(fm, tm)
    err = tm - fm
    rmin = float(np.amin(ratio))
    rmax = float(np.amax(ratio))
    return {
        "efficiency": float(efficiency),
        "uniformity": float(1 - (rmax - rmin) / (rmax + rmin)),
        "pkpk_err": float(err.size * float(np.amax(err) - np.amin(err))),
        "std_err": float(err.size * float(np.std(err))),
    }


class Hologram:
    """
    GPU Gerchberg-Saxton / weighted-GS phase retrieval for one padded complex field.
    Mirrors ``slmsuite.holography.algorithms.Hologram`` (_hologram.py:26-478).

    Parameters follow the reference constructor (_hologram.py:196-205):
    ``Hologram(target, amp=None, phase=None, slm_shape=None, dtype=np.float32,
    propagation_kernel=None, **kwargs)``.  ``device`` selects the CUDA device.
    """

    def __init__(self, target, amp=None, phase=None, slm_shape=None, dtype=np.float32,
                 propagation_kernel=None, device=0, **kwargs):
        # 1) shape voting, _hologram.py:296-356 (arrays / tuples; SLM objects are hardware scope)
        amp_shape = (np.nan, np.nan) if amp is None else np.shape(amp)
        phase_shape = (np.nan, np.nan) if phase is None else np.shape(phase)
        if slm_shape is None:
            slm_shape = (np.nan, np.nan)
        else:
            if hasattr(slm_shape, "slm"):  # CameraSLM-like
                slm_shape = slm_shape.slm.shape
            elif hasattr(slm_shape, "shape") and not isinstance(slm_shape, np.ndarray):
                slm_shape = slm_shape.shape
            if len(slm_shape) != 2:
                slm_shape = (np.nan, np.nan)
        stack = np.vstack((amp_shape, phase_shape, slm_shape)).astype(float)
        if np.all(np.isnan(stack)):
            self.slm_shape = None
        else:
            self.slm_shape = np.rint(np.nanmean(stack, axis=0)).astype(int)
            if not np.isnan(stack[0][0]) and not np.all(self.slm_shape == np.array(amp_shape)):
                raise ValueError(
                    "The shape of amplitude (via `amp` or SLM) is not equal to the "
                    "shapes of the provided initial phase (`phase`) or SLM (via `target` or `slm_shape`)"
                )
            if not np.isnan(stack[1][0]) and not np.all(self.slm_shape == np.array(phase_shape)):
                raise ValueError(
                    "The shape of the initial phase (`phase`) is not equal to the "
                    "shapes of the provided amplitude (via `amp` or SLM) or SLM (via `target` or `slm_shape`)"
                )
            if not np.isnan(stack[2][0]) and not np.all(self.slm_shape == np.array(slm_shape)):
                raise ValueError(
                    "The shape of SLM (via `target` or `slm_shape`) is not equal to the "
                    "shapes of the provided initial phase (`phase`) or amplitude (via `amp` or SLM)"
                )
            self.slm_shape = tuple(int(s) for s in self.slm_shape)

        # 1.5) target / shape, _hologram.py:358-387
        if target is None:
            raise ValueError("SLM shape must be provided through cameraslm=")
        if len(target) == 2 and np.ndim(target) == 1:
            self.shape = (int(target[0]), int(target[1]))
            target = None
        elif len(np.shape(target)) == 2:
            self.shape = tuple(int(s) for s in np.shape(target))
        else:
            raise ValueError(f"Unexpected target {target}.")
        if self.slm_shape is None:
            self.slm_shape = self.shape

        # 2) dtype, _hologram.py:391-398
        if dtype(0).nbytes == 4:
            self.dtype = np.float32
            self.dtype_complex = np.complex64
        elif dtype(0).nbytes == 8:
            raise ValueError("Data type float64/complex128 is not supported by the B200 path (float32/complex64 only).")
        else:
            raise ValueError(f"Data type {dtype} not supported.")

        for n in self.shape:
            if n < 16 or n > 8192 or (n & (n - 1)):
                raise ValueError(
                    f"Hologram shape {self.shape} must be powers of two in [16, 8192] per dimension on the B200 "
                    "path; use Hologram.get_padded_shape() to pad."
                )
        unpad_bounds(self.shape, self.slm_shape)  # raises if the SLM does not fit

        self._device = int(device)
        self._ctx = C.c_void_p()
        self._lib = _lib.lib()
        _lib.check(None, self._lib.slmgs_create(C.byref(self._ctx), self._device, self._batch_size(),
                                                self.shape[0], self.shape[1], self.slm_shape[0], self.slm_shape[1]))

        # amplitude, _hologram.py:401-405
        if amp is None:
            self._amp = 1 / np.sqrt(np.prod(self.slm_shape))
            self._check(self._lib.slmgs_set_amp_scalar(self._ctx, float(self._amp)))
        else:
            a = np.array(amp, dtype=self.dtype)
            a *= 1 / _norm(a)
            self._amp = a
            self._check(self._lib.slmgs_set_amp_array(self._ctx, _lib.fptr(_lib.f32(a)), 0))

        # propagation kernel, _hologram.py:408-415
        if propagation_kernel is None:
            self.propagation_kernel = None
        else:
            pk = np.array(propagation_kernel, dtype=self.dtype)
            if pk.shape != tuple(self.slm_shape):
                raise ValueError("Expected the propagation kernel to be the same shape as the SLM.")
            self.propagation_kernel = pk
            self._check(self._lib.slmgs_set_propagation(self._ctx, _lib.fptr(_lib.f32(pk))))

        self.flags = kwargs
        self._target = None
        self._mraf_cache = None
        self._zero_weights_active = False
        self._set_target(target, reset_weights=False)

        self._phase_set = False
        self.reset_phase(phase)
        self.reset(reset_phase=False, reset_flags=False)

    # ------------------------------------------------------------------ plumbing
    def _batch_size(self):
        return 1

    def _check(self, status):
        _lib.check(self._ctx, status)

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None and self._ctx.value:
                self._lib.slmgs_destroy(self._ctx)
                self._ctx = C.c_void_p()
        except Exception:
            pass

    def _bshape(self, shape):
        """Array shape of a per-hologram quantity (HologramBatch prepends the batch axis)."""
        return tuple(shape)

    def _download(self, fn, shape, dtype=np.float32):
        out = np.empty(self._bshape(shape), dtype=dtype)
        self._check(fn(self._ctx, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ------------------------------------------------------------------ state properties
    @property
    def amp(self):
        """Near-field source amplitude: scalar (np.float64) or L2-normalised array (_hologram.py:401-405)."""
        return self._amp

    @amp.setter
    def amp(self, value):
        """The reference stores ``amp`` as a plain attribute that users (and ``SimulatedCamera._get_image_hw``,
        hardware/cameras/simulated.py:364) overwrite WITHOUT normalisation; the same here."""
        if np.ndim(value) == 0:
            self._amp = value
            self._check(self._lib.slmgs_set_amp_scalar(self._ctx, float(value)))
        else:
            a = np.array(value, dtype=self.dtype)
            if a.shape != tuple(self.slm_shape):
                raise ValueError(f"amp of shape {a.shape} is not of slm_shape {self.slm_shape}")
            self._amp = a
            self._check(self._lib.slmgs_set_amp_array(self._ctx, _lib.fptr(_lib.f32(a)), 0))

    @property
    def phase(self):
        """Near-field phase, shape ``slm_shape`` (downloaded)."""
        return self._download(self._lib.slmgs_get_phase, self.slm_shape)

    @phase.setter
    def phase(self, value):
        self.reset_phase(value)

    @property
    def target(self):
        """Normalised far-field target amplitude, NaN marks the MRAF noise region (host copy)."""
        return self._target

    @property
    def weights(self):
        return self._download(self._lib.slmgs_get_weights, self.shape)

    @weights.setter
    def weights(self, value):
        self.set_weights(np.asarray(value))

    @property
    def amp_ff(self):
        """|farfield| of the last forward transform, or None before any (_hologram.py:472-473)."""
        if not self._amp_ff_set:
            return None
        return self._download(self._lib.slmgs_get_amp_ff, self.shape)

    @property
    def phase_ff(self):
        if self._phase_ff_none:
            return None
        return self._download(self._lib.slmgs_get_phase_ff, self.shape)

    @phase_ff.setter
    def phase_ff(self, value):
        if value is None:
            self._phase_ff_none = True
            return
        v = _lib.f32(value)
        if v.shape != tuple(self.shape):
            raise ValueError(f"phase_ff {v.shape} does not match shape {self.shape}")
        self._check(self._lib.slmgs_set_phase_ff(self._ctx, _lib.fptr(v)))
        self._phase_ff_none = False

    @property
    def farfield(self):
        """Complex far field of the current phase, ortho-normalised, centred (fftshift-ed) convention."""
        return self._download(self._lib.slmgs_get_farfield, self.shape, np.complex64)

    @property
    def nearfield(self):
        """Padded complex near field amp * exp(i phase) (_hologram.py:1000-1011), rebuilt on the host."""
        i0, i1, i2, i3 = unpad_bounds(self.shape, self.slm_shape)
        nf = np.zeros(self.shape, dtype=self.dtype_complex)
        ph = self.phase
        if self.propagation_kernel is not None:
            ph = ph + self.propagation_kernel
        nf[i0:i1, i2:i3] = self._amp * np.exp(1j * ph)
        return nf

    # ------------------------------------------------------------------ construction helpers
    def _set_target(self, new_target, reset_weights=False):
        """_hologram.py:741-766."""
        if new_target is None:
            self._target = np.zeros(shape=self.shape, dtype=self.dtype)
        else:
            t = np.array(new_target, dtype=self.dtype)
            if t.shape != tuple(self.shape):
                raise ValueError(f"Target shape {t.shape} does not match hologram shape {self.shape}")
            np.abs(t, out=t)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                with np.errstate(all="ignore"):
                    t *= 1 / _norm(t)
            self._target = t
        self._upload_target()
        if reset_weights:
            self.reset_weights()

    def _upload_target(self):
        self._mraf_cache = None
        self._check(self._lib.slmgs_set_target(self._ctx, _lib.fptr(_lib.f32(self._target)), 0))

    def set_target(self, new_target, reset_weights=False):
        """_hologram.py:768-784."""
        self._set_target(new_target=new_target, reset_weights=reset_weights)

    def _get_random_phase(self):
        """_hologram.py:528-534 (NumPy branch)."""
        rng = np.random.default_rng()
        return rng.uniform(-np.pi, np.pi, self.slm_shape).astype(self.dtype)

    def reset_phase(self, custom_phase=None, random_phase=None, quadratic_phase=None):
        """_hologram.py:536-601.  The analytic ``quadratic_phase`` preconditioner is host-side setup
        of the reference's toolbox (out of this path's scope) and raises if requested."""
        if custom_phase is not None:
            p = np.array(custom_phase, dtype=self.dtype)
            if not np.all(np.array(self.slm_shape) == np.array(p.shape)):
                raise ValueError(f"Reset phase of shape {p.shape} is not of slm_shape {self.slm_shape}")
        else:
            if quadratic_phase is None:
                quadratic_phase = self.flags.get("quadratic_phase", False)
            if random_phase is None:
                random_phase = self.flags.get("random_phase", 1)
            p = np.zeros(self.slm_shape, dtype=self.dtype)
            if quadratic_phase:
                raise NotImplementedError("quadratic_phase preconditioning is outside the GS/WGS hot path; pass phase=")
            if random_phase:
                p += random_phase * self._get_random_phase()
        self._check(self._lib.slmgs_set_phase(self._ctx, _lib.fptr(_lib.f32(p))))
        self._phase_set = True

    def reset_weights(self):
        """_hologram.py:603-614."""
        self._check(self._lib.slmgs_reset_weights(self._ctx))

    def reset(self, reset_phase=True, reset_flags=False):
        """_hologram.py:442-478."""
        if not self._phase_set or reset_phase:
            self.reset_phase()
        self.reset_weights()
        self.iter = 0
        self.stats = {"method": [], "flags": {}, "stats": {}}
        if reset_flags:
            self.flags = {"method": ""}
        self._amp_ff_set = False
        self._phase_ff_none = True

    @staticmethod
    def get_padded_shape(slm_shape, padding_order=1, square_padding=True, precision=np.inf,
                         precision_basis="kxy"):
        """_hologram.py:616-725 for the default ``precision=inf`` (finite precision needs a CameraSLM)."""
        if hasattr(slm_shape, "slm"):
            slm_shape = slm_shape.slm.shape
        elif hasattr(slm_shape, "shape") and not isinstance(slm_shape, (tuple, list, np.ndarray)):
            slm_shape = slm_shape.shape
        if np.isfinite(precision):
            raise NotImplementedError("precision-based padding needs a CameraSLM (hardware scope)")
        if padding_order > 0:
            padding_shape = np.power(2, np.ceil(np.log2(slm_shape)) + padding_order - 1).astype(int)
        else:
            padding_shape = np.array(slm_shape)
        shape = tuple(int(s) for s in np.max(np.vstack((padding_shape,)), axis=0))
        if square_padding:
            largest = max(shape)
            shape = (largest, largest)
        return shape

    # ------------------------------------------------------------------ accessors
    def get_phase(self, include_propagation=False):
        """_hologram.py:786-811: host phase + pi (or phase + kernel, without the pi, as the reference does)."""
        if include_propagation and self.propagation_kernel is not None:
            return self.phase + self.propagation_kernel
        return self.phase + np.pi

    # north_star names the accessor extract_phase(); the reference's is get_phase() (SURVEY.md note)
    extract_phase = get_phase

    def get_phase_gray(self, bitdepth=8, phase_correction=None):
        """
        The integer image ``SLM.set_phase(self.get_phase())`` would display on an SLM of ``bitdepth`` bits
        (hardware/slms/slm.py:636-690 and ``_phase2gray`` :695-743, ``phase_scaling == 1``), computed on the
        device and downloaded as uint8 (``bitdepth <= 8``) or uint16 -- a quarter / half of the bytes of
        ``get_phase()``.  ``phase_correction`` is the SLM's ``source["phase"]`` (added when ``phase_correct``).
        The result can be passed to the reference's ``SLM.set_phase``, which accepts integer data as is.
        """
        bitdepth = int(bitdepth)
        if bitdepth < 1 or bitdepth > 16:
            raise ValueError(f"bitdepth {bitdepth} not supported (1..16)")
        corr = None
        if phase_correction is not None:
            corr = np.ascontiguousarray(phase_correction, dtype=np.float64)
            if corr.shape != tuple(self.slm_shape):
                raise ValueError(f"phase_correction of shape {corr.shape} is not of slm_shape {self.slm_shape}")
        out = np.empty(self._bshape(self.slm_shape), dtype=np.uint8 if bitdepth <= 8 else np.uint16)
        self._check(self._lib.slmgs_get_phase_gray(
            self._ctx, bitdepth, None if corr is None else _lib.dptr(corr), out.ctypes.data_as(C.c_void_p)))
        return out

    def set_sparse(self, enabled=True):
        """
        Not in the reference.  Switches the sparse far-field path of the fused loop (default: automatic).  When the
        weights are zero on whole column tiles -- spot targets -- those tiles cannot contribute to
        ``farfield = weights * exp(i phase_ff)`` (_hologram.py:1601-1605), so the column kernels skip them and the
        row kernels do not move their columns.  Results are identical to the dense loop.
        """
        self._check(self._lib.slmgs_set_sparse(self._ctx, 1 if enabled else 0))

    def sparse_info(self):
        """(last fused run used the sparse path, active column tiles, column tiles)."""
        out = np.zeros(3, dtype=np.int32)
        self._check(self._lib.slmgs_sparse_info(self._ctx, _lib.iptr(out)))
        return bool(out[0]), int(out[1]), int(out[2])

    def get_amp(self):
        """_hologram.py:813-826."""
        return self._amp

    def set_weights(self, new_weights):
        """_hologram.py:828-840."""
        if tuple(np.shape(new_weights)) != tuple(self.shape):
            raise ValueError(f"New weights {np.shape(new_weights)} do not match target shape {self.shape}")
        self._check(self._lib.slmgs_set_weights(self._ctx, _lib.fptr(_lib.f32(new_weights))))

    def get_weights(self):
        """_hologram.py:842-850."""
        return self.weights

    def get_target(self):
        return self._target

    def get_farfield(self, shape=None, propagation_kernel=None, affine=None, get=True):
        """
        _hologram.py:853-931: complex DFT far field of the current phase, optionally on a different padded
        ``shape`` (changes the far-field resolution), with a different ``propagation_kernel`` (other depth), and
        resampled by an ``affine`` transform ``{"M", "b"}``.  The transform runs on the device (a temporary context
        for a different shape / kernel); the cubic affine resampling is the same SciPy call
        the reference's NumPy backend makes.  Like the reference, refreshes ``amp_ff`` when the shape matches.
        """
        if shape is None:
            shape = self.shape
        if len(shape) == 1:
            shape = self.slm_shape
        shape = tuple(int(s) for s in shape)
        own_kernel = propagation_kernel is None
        if own_kernel:
            propagation_kernel = self.propagation_kernel
        if shape == tuple(self.shape) and own_kernel:
            ff = self.farfield
            self._amp_ff_set = True
        else:
            if np.isscalar(propagation_kernel) or (propagation_kernel is not None and np.ndim(propagation_kernel) == 0):
                # "Zeroing can force no kernel to be applied and yield the raw DFT" (:871-873)
                propagation_kernel = None if float(propagation_kernel) == 0 else np.full(
                    self.slm_shape, float(propagation_kernel), dtype=self.dtype)
            if self._batch_size() != 1:
                raise NotImplementedError("get_farfield(shape=, propagation_kernel=) is per hologram")
            tmp = Hologram(shape, amp=None if np.isscalar(self._amp) else self._amp, phase=self.phase,
                           slm_shape=self.slm_shape, propagation_kernel=propagation_kernel, device=self._device)
            if np.isscalar(self._amp):
                tmp._check(tmp._lib.slmgs_set_amp_scalar(tmp._ctx, float(self._amp)))
            ff = tmp.farfield
            del tmp
        if affine is not None:
            from scipy.ndimage import affine_transform as sp_affine_transform

            sp_affine_transform(input=ff, matrix=affine["M"], offset=affine["b"], output_shape=shape, order=3,
                                output=ff, mode="constant", cval=0)
        return ff

    # ------------------------------------------------------------------ optimisation
    def optimize(self, method="GS", maxiter=20, verbose=True, callback=None, feedback=None,
                 stat_groups=[], **kwargs):
        """_hologram.py:1076-1368."""
        name = kwargs.pop("name", None)
        self._update_flags(method, verbose, feedback, stat_groups, **kwargs)
        iterations = range(maxiter)
        if verbose and maxiter > 1:
            try:
                from tqdm.auto import tqdm
                iterations = tqdm(iterations, desc=name)
            except Exception:
                pass
        if "GS" in method:
            self.optimize_gs(iterations, callback)
        else:
            raise ValueError(f"Unsupported optimization method '{method}'")

    def _update_flags(self, method, verbose, feedback, stat_groups, **kwargs):
        """_hologram.py:1370-1424."""
        methods = list(ALGORITHM_DEFAULTS.keys())
        if method not in methods:
            raise ValueError(
                "Unrecognized method '{}'.\n"
                "Valid methods include {}".format(method, methods)
            )
        self.flags["method"] = method
        for flag, value in ALGORITHM_DEFAULTS[method].items():
            if flag not in self.flags:
                self.flags[flag] = value
        if "fixed_phase" not in self.flags:
            self.flags["fixed_phase"] = False
        for flag in kwargs:
            self.flags[flag] = kwargs[flag]
        for group in stat_groups:
            if group not in FEEDBACK_OPTIONS:
                raise ValueError(
                    "Statistics group '{}' not recognized as a feedback option.\n"
                    "Valid options: {}".format(group, FEEDBACK_OPTIONS)
                )
        self.flags["stat_groups"] = stat_groups
        if feedback is not None:
            if feedback not in FEEDBACK_OPTIONS:
                raise ValueError(
                    "Feedback '{}' not recognized as a feedback option.\n"
                    "Valid options: {}".format(feedback, FEEDBACK_OPTIONS)
                )
            self.flags["feedback"] = feedback
        if verbose > 1:
            import pprint
            print(f"Optimizing with '{method}' using the following method-specific flags:")
            pprint.pprint({k: v for (k, v) in self.flags.items() if k in ALGORITHM_DEFAULTS[method]})
            print("", end="", flush=True)

    def _mraf_enabled(self):
        """_hologram.py:1495-1501 (``isnan(sum(target))``), evaluated once per target: the reference sums the
        full target on every optimize() call, which costs milliseconds on the host at 4096^2."""
        if self._mraf_cache is None:
            with np.errstate(all="ignore"):
                self._mraf_cache = bool(np.isnan(np.sum(self._target)))
        return self._mraf_cache

    def _zero_weights_on(self, mraf):
        """The MRAF zero-region accumulator exists once ``zero_factor`` was non-zero with a non-empty zero region and is
        used from then on (``hasattr(self, "zero_weights")``, _hologram.py:1511-1515, :1613-1616)."""
        if mraf and not self._zero_weights_active and self.flags.get("zero_factor", 0) != 0:
            with np.errstate(invalid="ignore"):
                self._zero_weights_active = bool(np.any(self._target == 0))
        return mraf and self._zero_weights_active

    def _fusable(self, callback):
        """
        The fused launch sequence (``slmgs_run``) is legal when nothing on the host needs the far field
        between the transforms (SURVEY.md 7 "callback contract"): no callback, no statistics, device-side
        feedback, and no efficiency-triggered Kim fixing (needs statistics).
        """
        fl = self.flags
        if callback is not None or len(fl["stat_groups"]) > 0:
            return False
        if fl.get("raw_stats", False):
            return False
        m = fl["method"]
        if m == "GS":
            return True
        if fl["feedback"] not in self._device_feedbacks():
            return False
        if m == "WGS-Kim" and fl.get("fix_phase_efficiency", None) is not None:
            return False
        return True

    def _device_feedbacks(self):
        return ("computational",)

    def _feedback_params(self):
        """(slmgs_params.feedback, slmgs_params.spot_width) for the current flags."""
        return 0, 0

    def _iteration_params(self, mraf, stepped):
        """
        Host part of ``_gs_farfield_routines`` (_hologram.py:1550-1605): decides whether the weights are
        updated this iteration and runs the WGS-Kim fixed-phase state machine on ``flags`` and the
        recorded flag history.  Returns the flags of this iteration for the device.
        """
        fl = self.flags
        method = fl["method"]
        update = ("WGS" in method) and self.iter > 0
        store = False
        if update:
            if "Kim" in method:
                was_not_fixed = not fl["fixed_phase"]
                if fl["fix_phase_efficiency"] is not None:
                    stats = self.stats["stats"]
                    groups = tuple(stats.keys())
                    if len(groups) == 0:
                        raise ValueError("Must track statistics to fix phase based on efficiency!")
                    eff = stats[groups[-1]]["efficiency"][self.iter]
                    if eff > fl["fix_phase_efficiency"]:
                        fl["fixed_phase"] = True
                if was_not_fixed and self.iter >= fl["fix_phase_iteration"] - 1:
                    hist = self.stats["flags"]["fixed_phase"]
                    if np.all([not hist[-1 - i] for i in range(fl["fix_phase_iteration"])]):
                        fl["fixed_phase"] = True
                if (fl["fixed_phase"] and self._phase_ff_none) or was_not_fixed:
                    store = True
            else:
                fl["fixed_phase"] = False
        if not fl.get("fixed_phase", False):
            mode = _lib.PHASE_COMPUTE_STORE if stepped else _lib.PHASE_COMPUTE
        elif self._phase_ff_none or store:
            mode = _lib.PHASE_COMPUTE_STORE  # the iteration that fixes keeps angle(farfield)
        else:
            mode = _lib.PHASE_STORED
        self._phase_ff_none = False
        mf = fl.get("mraf_factor", None)
        return _lib.Params(
            method=_lib.METHODS[method],
            update_weights=int(update),
            phase_mode=mode,
            feedback_exponent=float(fl.get("feedback_exponent", 0.0) or 0.0),
            feedback_factor=float(fl.get("feedback_factor", 0.0) or 0.0),
            mraf=int(mraf),
            mraf_has_factor=int(mf is not None),
            mraf_factor=float(mf if mf is not None else 1.0),
            feedback=self._feedback_params()[0],
            spot_width=self._feedback_params()[1],
            zero_weights=int(self._zero_weights_on(mraf)),
            zero_factor=float(fl.get("zero_factor", 1)),
        )

    def optimize_gs(self, iterations, callback):
        """_hologram.py:1427-1493."""
        mraf = self._mraf_enabled()
        if self._fusable(callback):
            # every host decision of the loop only depends on the iteration count: replay the
            # bookkeeping, then run all iterations in one asynchronous launch sequence
            plist = []
            for _ in iterations:
                self._update_stats(self.flags["stat_groups"])
                plist.append(self._iteration_params(mraf, stepped=False))
                self.iter += 1
            arr = (_lib.Params * max(len(plist), 1))(*plist)
            self._check(self._lib.slmgs_run(self._ctx, arr, len(plist), 1))
            self._amp_ff_set = True
            self._phase_ff_none = False
            return

        for _ in iterations:
            self._check(self._lib.slmgs_forward(self._ctx))  # (A)
            self._amp_ff_set = True
            if callback is not None:  # (B.1)
                if callback(self):
                    break
            self._update_stats(self.flags["stat_groups"])  # (B.2)
            params = self._iteration_params(mraf, stepped=True)  # (B.3)
            if params.update_weights:
                self._update_weights(params)
            self._check(self._lib.slmgs_constrain_inverse(self._ctx, C.byref(params)))  # (B.3) + (C)
            self.iter += 1
        self._check(self._lib.slmgs_populate(self._ctx))
        self._amp_ff_set = True
        self._phase_ff_none = False

    def _update_weights(self, params):
        """_hologram.py:1914-1922."""
        feedback = self.flags["feedback"]
        if feedback == "computational":
            self._check(self._lib.slmgs_update_weights(self._ctx, C.byref(params)))

    # ------------------------------------------------------------------ statistics
    def _stats_pixel(self):
        """``_calculate_stats(amp_ff, target)`` (_stats.py:7-116) from device-side reductions."""
        o8 = np.zeros((self._batch_size(), 8), dtype=np.float64)
        o2 = np.zeros((self._batch_size(), 2), dtype=np.float64)
        self._check(self._lib.slmgs_stats_pixel(self._ctx, _lib.dptr(o8), _lib.dptr(o2)))
        out = []
        for b in range(o8.shape[0]):
            fsum, tsum, ft, rmin, rmax, emin, emax, esum = o8[b]
            esq, cnt = o2[b]
            mean = esum / cnt
            var = max(esq / cnt - mean * mean, 0.0)
            out.append({
                "efficiency": float(ft * ft / (fsum * tsum)),
                "uniformity": float(1 - (rmax - rmin) / (rmax + rmin)),
                "pkpk_err": float(cnt * (emax - emin)),
                "std_err": float(cnt * np.sqrt(var)),
            })
        return out

    def _calculate_stats_computational(self, stats, stat_groups=[]):
        """_stats.py:118-128."""
        if "computational" in stat_groups:
            stats["computational"] = self._stats_pixel()[0]

    def _update_stats_dictionary(self, stats):
        """_stats.py:130-208 (``raw_stats`` keeps the far field like the reference)."""
        M = len(self.stats["method"])
        diff = self.iter + 1 - M
        if diff > 0:
            self.stats["method"].extend(["" for _ in range(diff)])
            M = self.iter + 1
        self.stats["method"][self.iter] = self.flags["method"]
        flaglist = set(self.flags.keys()).union(set(self.stats["flags"].keys()))
        for flag in flaglist:
            if flag not in self.stats["flags"]:
                self.stats["flags"][flag] = [np.nan for _ in range(M)]
            else:
                diff = self.iter + 1 - len(self.stats["flags"][flag])
                if diff > 0:
                    self.stats["flags"][flag].extend([np.nan for _ in range(diff)])
            if flag in self.flags:
                self.stats["flags"][flag][self.iter] = self.flags[flag]
        grouplist = set(stats.keys()).union(set(self.stats["stats"].keys()))
        if len(grouplist) > 0:
            statlists = [set(stats[group].keys()) for group in stats.keys()]
            if len(self.stats["stats"].keys()) > 0:
                key = next(iter(self.stats["stats"]))
                statlists.append(set(self.stats["stats"][key].keys()))
            statlist = set.union(*statlists)
            for group in grouplist:
                if group not in self.stats["stats"]:
                    self.stats["stats"][group] = {}
                for stat in statlist:
                    if stat not in self.stats["stats"][group]:
                        self.stats["stats"][group][stat] = [np.nan for _ in range(M)]
                    else:
                        diff = self.iter + 1 - len(self.stats["stats"][group][stat])
                        if diff > 0:
                            self.stats["stats"][group][stat].extend([np.nan for _ in range(diff)])
                    if group in stats.keys() and stat in stats[group].keys():
                        self.stats["stats"][group][stat][self.iter] = stats[group][stat]
        if self.flags.get("raw_stats", False):
            if "raw_farfield" not in self.stats:
                self.stats["raw_farfield"] = []
            diff = self.iter + 1 - len(self.stats["raw_farfield"])
            if diff > 0:
                self.stats["raw_farfield"].extend([np.nan for _ in range(diff)])
            self.stats["raw_farfield"][self.iter] = self.farfield

    def _update_stats(self, stat_groups=[]):
        """_stats.py:210-223."""
        stats = {}
        self._calculate_stats_computational(stats, stat_groups)
        self._update_stats_dictionary(stats)
