"""
``CompressedSpotHologram`` on the B200 ("next" row 4): host mirror of
``slmsuite.holography.algorithms.CompressedSpotHologram`` (_spots.py:178-1019).

Instead of a DFT grid every spot owns a phase kernel ``sum_d spot_zernike[d, n] Z_d(x, y)`` (free-floating k-vectors,
focus, any Zernike term), and the two maps of the GS / WGS loop are direct sums evaluated by hand-written kernels
(``slmsuite_b200/csrc/slmgs_compressed.h``) behind the C ABI (``slmgs_comp_*``, include/slmgs.h).  The loop itself --
flags, WGS-Kim state machine, weight update, MRAF via NaN / zero entries of ``spot_amp`` -- is the base ``Hologram``'s,
acting on N-vectors.

The reference needs a ``FourierSLM`` hardware object.  What it reads from it on this path is the SLM grid and the
aperture scaling, so either pass ``cameraslm=`` (anything with ``.slm.grid`` and ``.slm.get_source_zernike_scaling()``,
e.g. the reference's own object) or ``slm_grid=(x_grid, y_grid)`` + ``zernike_scaling=``.  Camera feedback
(``experimental_spot`` / ``external_spot``) and the ``"ij"`` basis need the camera and are outside this path.
"""
import ctypes as C
import warnings
from math import factorial

import numpy as np

from . import _lib
from .hologram import Hologram, _norm


# --------------------------------------------------------------------------- Zernike -> monomials (host setup)
def zernike_monomials(index):
    """{(a, b): coefficient} of x^a y^b for the Zernike polynomial of ANSI ``index`` (toolbox/phase.py:1357-1420,
    the expansion of doi:10.1117/12.294412; ANSI -> (n, l) as toolbox/phase.py:603-605)."""
    index = int(index)
    n = int(np.floor(0.5 * np.sqrt(8 * index + 1) - 0.5))
    l = -(2 * index - n * (n + 2))
    if l % 2:
        q = (abs(l) - 1) // 2
    elif l > 0:
        q = abs(l) // 2 - 1
    else:
        q = abs(l) // 2
    p = 0 if l <= 0 else 1
    l = abs(l)
    m = (n - l) // 2
    out = {}
    for i in range(q + 1):
        for j in range(m + 1):
            for k in range(m - j + 1):
                f = -1 if (i + j) % 2 else 1
                f *= factorial(l) / (factorial(2 * i + p) * factorial(l - 2 * i - p))
                f *= factorial(m - j) / (factorial(k) * factorial(m - j - k))
                f *= float(factorial(n - j)) / (factorial(j) * factorial(m - j) * factorial(n - m - j))
                key = (n - 2 * (i + j + k) - p, 2 * (i + k) + p)
                out[key] = out.get(key, 0) + int(f)
    return {k: v for k, v in out.items() if v != 0}


def zernike_indices_parse(D):
    """toolbox/phase.py:923-962 for ``indices=None``."""
    if D == 2:
        return np.array([2, 1])
    if D == 3:
        return np.array([2, 1, 4])
    if D == 4:
        return np.array([2, 1, 4, 3])
    return np.hstack((np.array([2, 1, 4, 3]), np.arange(5, D + 1)))


def monomial_table(zernike_basis):
    """(px[M], py[M], c[M][D]) with Z_d = sum_m c[m, d] x^px[m] y^py[m]."""
    terms = {}
    D = len(zernike_basis)
    for d, idx in enumerate(zernike_basis):
        if idx < 0:
            raise ValueError("special (negative) Zernike indices (vortex) are not supported on the B200 path")
        for key, coef in zernike_monomials(idx).items():
            terms.setdefault(key, np.zeros(D))[d] = coef
    keys = sorted(terms, key=lambda ab: (ab[0] + ab[1], ab[1]))
    px = np.array([k[0] for k in keys], dtype=np.int64)
    py = np.array([k[1] for k in keys], dtype=np.int64)
    return px, py, np.array([terms[k] for k in keys], dtype=np.float64).reshape(len(keys), D)


class CompressedSpotHologram(Hologram):
    """
    ``CompressedSpotHologram(spot_vectors, basis="kxy", spot_amp=None, cameraslm=None, cuda=False, **kwargs)``,
    _spots.py:214-221, plus ``slm_grid`` / ``zernike_scaling`` / ``amp`` / ``phase`` / ``device`` in place of the hardware
    object.  ``basis``: ``"kxy"`` (2 or 3 rows: x, y[, focal power]), ``"zernike"`` (default basis of that dimension)
    or a list of ANSI indices.  ``cuda`` is accepted and ignored (there is only the CUDA path).
    """

    def __init__(self, spot_vectors, basis="kxy", spot_amp=None, cameraslm=None, cuda=False, slm_grid=None,
                 zernike_scaling=None, amp=None, phase=None, device=0, **kwargs):
        if cameraslm is None and slm_grid is None:
            raise ValueError("cameraslm must be passed.")
        if cameraslm is not None:
            slm = cameraslm.slm if hasattr(cameraslm, "slm") else cameraslm
            slm_grid = slm.grid
            if zernike_scaling is None:
                zernike_scaling = slm.get_source_zernike_scaling()
            if amp is None and hasattr(slm, "_get_source_amplitude"):
                amp = slm._get_source_amplitude()  # _feedback.py:79
        if zernike_scaling is None:
            raise ValueError("zernike_scaling (slm.get_source_zernike_scaling()) is needed with slm_grid")
        spot_vectors = np.array(spot_vectors, dtype=float)
        if spot_vectors.ndim != 2:
            raise ValueError("spot_vectors must have shape (D, N)")
        D, N = spot_vectors.shape

        # _spots.py:346-356
        if spot_amp is not None:
            self.spot_amp = np.array(spot_amp)
            if self.spot_amp.size != N:
                raise ValueError(f"spot_amp (length {self.spot_amp.size}) must have the same length as the provided spots ({D}).")
        else:
            self.spot_amp = np.full(N, 1.0 / np.sqrt(N))

        # _spots.py:358-389
        if isinstance(basis, str):
            self.zernike_basis = zernike_indices_parse(D)
        else:
            self.zernike_basis = np.ravel(basis)
            basis = "zernike"
            if len(self.zernike_basis) != D:
                raise ValueError(f"zernike_basis (length {len(self.zernike_basis)}) must have the same dimension "
                                 f"as the provided spots ({D}).")
            if 0 in self.zernike_basis:
                warnings.warn("Found ANSI index '0' (Zernike piston) in the zernike_basis; "
                              "this is not necessary as spot phase is controlled externally.")
        if not np.any(self.zernike_basis == 2) or not np.any(self.zernike_basis == 1):
            raise ValueError("Compressed basis must include x, y (Zernike ANSI indices 2, 1)")

        # _spots.py:391-428 with toolbox.convert_vector (toolbox/__init__.py:312-316, :355-356, :390-391)
        scale = 2 * np.pi * np.reciprocal(float(zernike_scaling))
        if basis == "zernike":
            self.spot_zernike = np.array(spot_vectors)
            cart = [int(np.argwhere(self.zernike_basis == 2)[0][0]), int(np.argwhere(self.zernike_basis == 1)[0][0])]
            self.spot_kxy = self.spot_zernike[cart, :] / scale
        elif basis == "kxy":
            if D not in (2, 3):
                raise ValueError("basis 'kxy' expects 2 or 3 rows")
            self.spot_kxy = np.array(spot_vectors)
            self.spot_zernike = np.array(spot_vectors)
            self.spot_zernike[:2] = spot_vectors[:2] * scale
            if D == 3:
                self.spot_zernike[2] = spot_vectors[2] * ((scale * scale) / (8 * np.pi))
        else:
            raise NotImplementedError(f"basis '{basis}' needs camera hardware (outside the GS/WGS hot path)")
        self.spot_ij = None
        self.spot_integration_width_ij = None

        x_grid, y_grid = slm_grid
        self.slm_shape = tuple(int(s) for s in np.shape(x_grid))
        self.shape = self.slm_shape  # _spots.py:489
        self.dtype = np.float32
        self.dtype_complex = np.complex64
        self.cuda = True
        self.cameraslm = cameraslm
        self.propagation_kernel = None

        # host setup of the phase kernels: the basis functions Z_d on the aperture-scaled grid (_spots.py:609-614 stores
        # the scaled grid as complex64, i.e. rounded to float32), evaluated once in float64 from their monomial
        # expansion; the device then needs D double FMAs per (pixel, spot) pair
        px, py, c = monomial_table(self.zernike_basis)
        self._px, self._py, self._c = px, py, c
        x = np.array(np.asarray(x_grid) * zernike_scaling, dtype=np.float32).astype(np.float64).ravel()
        y = np.array(np.asarray(y_grid) * zernike_scaling, dtype=np.float32).astype(np.float64).ravel()
        if D > 10:
            raise ValueError("a Zernike basis of more than 10 terms is not supported on the B200 path")
        mono = np.zeros((D, x.size), dtype=np.float64)
        for m in range(len(px)):
            term = x ** int(px[m]) * y ** int(py[m])
            for d in range(D):
                if c[m, d] != 0:
                    mono[d] += c[m, d] * term

        self._device = int(device)
        self._ctx = C.c_void_p()
        self._lib = _lib.lib()
        status = self._lib.slmgs_comp_create(C.byref(self._ctx), self._device, self.slm_shape[0], self.slm_shape[1],
                                             N, D)
        if status != _lib.OK:
            msg = self._lib.slmgs_comp_last_error(None)
            raise (ValueError if status == _lib.ERR_INVALID else RuntimeError)(msg.decode() if msg else "slmgs error")
        self._mono = np.ascontiguousarray(mono)
        self._upload_basis()

        # amplitude, _hologram.py:401-405
        if amp is None:
            self._amp = 1 / np.sqrt(np.prod(self.slm_shape))
            self._check(self._lib.slmgs_comp_set_amp_scalar(self._ctx, float(self._amp)))
        else:
            a = np.array(amp, dtype=self.dtype)
            if a.shape != self.slm_shape:
                raise ValueError("The shape of amplitude is not equal to the shape of the SLM")
            a *= 1 / _norm(a)
            self._amp = a
            self._check(self._lib.slmgs_comp_set_amp_array(self._ctx, _lib.fptr(_lib.f32(a))))

        self.flags = kwargs
        self._target = None
        self._mraf_cache = None
        self._zero_weights_active = False
        self.set_target(new_target=self.spot_amp, reset_weights=False)
        self._phase_set = False
        self.reset_phase(phase)
        self.reset(reset_phase=False, reset_flags=False)
        self.external_spot_amp = np.ones(self._target.shape)

    # ------------------------------------------------------------------ plumbing
    def _check(self, status):
        if status == _lib.OK:
            return
        msg = self._lib.slmgs_comp_last_error(self._ctx)
        msg = msg.decode() if msg else "slmgs error {}".format(status)
        if status == _lib.ERR_INVALID:
            raise ValueError(msg)
        if status == _lib.ERR_OOM:
            raise MemoryError(msg)
        raise _lib.SlmgsError(msg)

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                self._lib.slmgs_comp_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def __len__(self):
        """_spots.py:547-556."""
        return self.spot_amp.size

    def _upload_basis(self):
        cw = np.ascontiguousarray(self.spot_zernike, dtype=np.float64)  # (D, N): weights of the basis functions
        self._check(self._lib.slmgs_comp_set_basis(self._ctx, _lib.dptr(self._mono), _lib.dptr(cw)))
        self._spot_zernike_cached = self.spot_zernike.copy()

    def _check_spot_zernike_change(self):
        """_spots.py:638-650: the user may move the spots between optimize() calls."""
        if np.any(self._spot_zernike_cached != self.spot_zernike):
            self._upload_basis()

    def _vec(self, fn, dtype=np.float32):
        out = np.empty(len(self), dtype=dtype)
        self._check(fn(self._ctx, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    @staticmethod
    def get_padded_shape(*args, **kwargs):
        """_spots.py:558-564."""
        raise NameError("CompressedSpotHologram does not use a DFT grid and does not need padding.")

    # ------------------------------------------------------------------ state
    @property
    def amp(self):
        return self._amp

    @property
    def phase(self):
        out = np.empty(self.slm_shape, dtype=np.float32)
        self._check(self._lib.slmgs_comp_get_phase(self._ctx, _lib.fptr(out)))
        return out

    @phase.setter
    def phase(self, value):
        self.reset_phase(value)

    @property
    def weights(self):
        return self._vec(self._lib.slmgs_comp_get_weights)

    @weights.setter
    def weights(self, value):
        self.set_weights(np.asarray(value))

    @property
    def amp_ff(self):
        return self._vec(self._lib.slmgs_comp_get_amp_ff) if self._amp_ff_set else None

    @property
    def phase_ff(self):
        return None if self._phase_ff_none else self._vec(self._lib.slmgs_comp_get_phase_ff)

    @phase_ff.setter
    def phase_ff(self, value):
        if value is None:
            self._phase_ff_none = True
            return
        v = _lib.f32(np.ravel(value))
        if v.size != len(self):
            raise ValueError("phase_ff must have one entry per spot")
        self._check(self._lib.slmgs_comp_set_phase_ff(self._ctx, _lib.fptr(v)))
        self._phase_ff_none = False

    @property
    def farfield(self):
        """Normalised complex spot amplitudes (what the reference holds after ``_nearfield2farfield``)."""
        if not self._amp_ff_set:
            self._check(self._lib.slmgs_comp_forward(self._ctx, 0))
            self._amp_ff_set = True
        return self._vec(self._lib.slmgs_comp_get_farfield, np.complex64)

    @property
    def nearfield(self):
        return (self._amp * np.exp(1j * self.phase)).astype(self.dtype_complex)

    def set_target(self, new_target=None, reset_weights=False):
        """_spots.py:917-948."""
        if new_target is None:
            t = np.array(self.spot_amp, dtype=self.dtype)
        else:
            new_target = np.squeeze(np.ravel(new_target))
            if new_target.shape != (len(self),):
                raise ValueError("Target must be of appropriate shape. "
                                 "Initialize a new Hologram if a different shape is desired.")
            t = np.array(new_target, dtype=self.dtype)
            self.spot_amp = np.array(new_target, dtype=self.dtype)
        np.abs(t, out=t)
        with np.errstate(all="ignore"):
            t *= 1 / _norm(t)
        self._target = t
        self._mraf_cache = None
        self._check(self._lib.slmgs_comp_set_target(self._ctx, _lib.fptr(_lib.f32(t))))
        if reset_weights:
            self.reset_weights()

    def reset_phase(self, custom_phase=None, random_phase=None, quadratic_phase=None):
        """_hologram.py:536-601."""
        if custom_phase is not None:
            p = np.array(custom_phase, dtype=self.dtype)
            if tuple(p.shape) != tuple(self.slm_shape):
                raise ValueError(f"Reset phase of shape {p.shape} is not of slm_shape {self.slm_shape}")
        else:
            if quadratic_phase or self.flags.get("quadratic_phase", False):
                raise NotImplementedError("quadratic_phase preconditioning is outside the GS/WGS hot path; pass phase=")
            if random_phase is None:
                random_phase = self.flags.get("random_phase", 1)
            p = np.zeros(self.slm_shape, dtype=self.dtype)
            if random_phase:
                p += random_phase * self._get_random_phase()
        self._check(self._lib.slmgs_comp_set_phase(self._ctx, _lib.fptr(_lib.f32(p))))
        self._phase_set = True
        self._amp_ff_set = False

    def reset_weights(self):
        """_hologram.py:603-614: weights = nan_to_num(target, nan=0)."""
        w = np.nan_to_num(self._target.copy(), nan=0)
        self._check(self._lib.slmgs_comp_set_weights(self._ctx, _lib.fptr(_lib.f32(w))))

    def set_weights(self, new_weights):
        w = _lib.f32(np.ravel(new_weights))
        if w.size != len(self):
            raise ValueError(f"New weights {np.shape(new_weights)} do not match the number of spots {len(self)}")
        self._check(self._lib.slmgs_comp_set_weights(self._ctx, _lib.fptr(w)))

    # things of the grid-based Hologram that make no sense here
    def get_phase_gray(self, bitdepth=8, phase_correction=None):
        raise NotImplementedError("get_phase_gray is implemented for the grid-based holograms only")

    def get_farfield(self, *args, **kwargs):
        raise NotImplementedError("CompressedSpotHologram has no DFT grid; use .farfield (complex spot amplitudes)")

    def set_sparse(self, enabled=True):
        pass

    def sparse_info(self):
        return False, len(self), len(self)

    # ------------------------------------------------------------------ loop
    def _zero_weights_on(self, mraf):
        if mraf and self.flags.get("zero_factor", 0) != 0:
            raise NotImplementedError("the MRAF zero_factor accumulator is not supported for compressed holograms")
        return False

    def _check_feedback(self):
        """_spots.py:950-966."""
        fb = self.flags.get("feedback", "computational")
        if fb == "computational":
            fb = self.flags["feedback"] = "computational_spot"
        if fb == "experimental":
            warnings.warn("CompressedSpotHologram feedback 'experimental' is interpreted as 'experimental_spot'")
            fb = self.flags["feedback"] = "experimental_spot"
        if fb in ("experimental_spot", "external_spot"):
            raise NotImplementedError(f"Feedback '{fb}' needs camera hardware (outside the GS/WGS hot path)")
        if fb != "computational_spot":
            raise ValueError("Feedback '{}' not recognized.".format(fb))

    def _update_stats(self, stat_groups=[]):
        """_spots.py:1004-1019: only the experimental group is computed by the reference; bookkeeping only here."""
        self._update_stats_dictionary({})

    def optimize_gs(self, iterations, callback):
        """_hologram.py:1427-1493 with the compressed maps."""
        self._check_spot_zernike_change()
        mraf = self._mraf_enabled()
        if "WGS" in self.flags["method"]:
            self._check_feedback()
        # the fused replay is legal when nothing on the host reads the far field between the maps (as Hologram._fusable):
        # raw_stats / statistics groups need the far field of every iteration
        if callback is None and not self.flags.get("raw_stats", False) and len(self.flags["stat_groups"]) == 0:
            plist = []
            for _ in iterations:
                self._update_stats(self.flags["stat_groups"])
                plist.append(self._iteration_params(mraf, stepped=False))
                self.iter += 1
            arr = (_lib.Params * max(len(plist), 1))(*plist)
            self._check(self._lib.slmgs_comp_run(self._ctx, arr, len(plist), 1))
        else:
            for _ in iterations:
                self._check(self._lib.slmgs_comp_forward(self._ctx, 0))  # so the callback sees farfield / amp_ff
                self._amp_ff_set = True
                if callback is not None and callback(self):
                    break
                self._update_stats(self.flags["stat_groups"])
                params = self._iteration_params(mraf, stepped=True)
                self._check(self._lib.slmgs_comp_run(self._ctx, C.byref(params), 1, 0))
                self.iter += 1
            self._check(self._lib.slmgs_comp_forward(self._ctx, 1))
        self._amp_ff_set = True
        self._phase_ff_none = False


# --------------------------------------------------------------------------- one hologram on several GPUs
class ShardedCompressedSpotHologram(CompressedSpotHologram):
    """
    ONE compressed spot hologram spread over several GPUs (not in the reference, whose kernels are single-GPU).

    The SLM rows are split into contiguous slabs, one per rank; every rank knows all N spots.  near -> far is a sum
    over pixels, so after the local pass the ranks all-reduce the N complex accumulators (16 N bytes -- the only
    exchange of an iteration), run the identical N-vector stage (normalisation, WGS update, WGS-Kim phase, MRAF) and
    project their own slab.  ``phase`` / ``get_phase()`` gather the slabs.  Same constructor as
    ``CompressedSpotHologram`` plus ``comm`` (default: ``slmsuite_b200.comm.default()``: NCCL behind the C ABI on GPUs;
    any object with the same ``rank`` / ``world`` / ``allreduce_f64`` / ``allgather_rows`` interface works).
    """

    def __init__(self, spot_vectors, basis="kxy", spot_amp=None, cameraslm=None, cuda=False, slm_grid=None,
                 zernike_scaling=None, amp=None, phase=None, device=0, comm=None, **kwargs):
        if comm is None:
            from . import comm as _comm

            comm = _comm.default()
        self._comm = comm
        rank, world = self._comm.rank, self._comm.world
        if cameraslm is not None:
            slm = cameraslm.slm if hasattr(cameraslm, "slm") else cameraslm
            slm_grid = slm.grid
            if zernike_scaling is None:
                zernike_scaling = slm.get_source_zernike_scaling()
            if amp is None and hasattr(slm, "_get_source_amplitude"):
                amp = slm._get_source_amplitude()
        if slm_grid is None:
            raise ValueError("cameraslm must be passed.")
        x_grid, y_grid = (np.asarray(g) for g in slm_grid)
        full = tuple(int(s) for s in x_grid.shape)
        bounds = [(full[0] * r) // world for r in range(world + 1)]
        self._rows = [bounds[r + 1] - bounds[r] for r in range(world)]
        if min(self._rows) < 1:
            raise ValueError("more ranks than SLM rows")
        r0, r1 = bounds[rank], bounds[rank + 1]
        self._r0, self._r1, self._full_shape = r0, r1, full
        if amp is not None:  # the reference normalises over the whole SLM (_hologram.py:404-405): do it before slicing
            amp = np.array(amp, dtype=np.float32)
            amp *= 1 / _norm(amp)
            amp = np.ascontiguousarray(amp[r0:r1])
        if phase is not None:
            phase = np.asarray(phase)[r0:r1]
        super().__init__(spot_vectors, basis=basis, spot_amp=spot_amp, cameraslm=None, cuda=cuda,
                         slm_grid=(x_grid[r0:r1], y_grid[r0:r1]), zernike_scaling=zernike_scaling, amp=amp, phase=phase,
                         device=device, **kwargs)
        self.cameraslm = cameraslm
        if amp is None:
            # scalar amplitude 1/sqrt(h w) of the WHOLE SLM (the constructor took the slab's size)
            self._amp = 1 / np.sqrt(np.prod(full))
            self._check(self._lib.slmgs_comp_set_amp_scalar(self._ctx, float(self._amp)))
        else:
            # undo the slab-local renormalisation of the base constructor
            self._amp = amp
            self._check(self._lib.slmgs_comp_set_amp_array(self._ctx, _lib.fptr(_lib.f32(amp))))
        self._local_shape = tuple(self.slm_shape)
        self.slm_shape = self.shape = full  # the public shape is the whole SLM; the device holds this rank's slab
        self._on_device = _lib.library_path() == _lib.DEFAULT_LIBRARY

    # the slab is what the device holds; the public shape is the whole SLM
    @property
    def phase(self):
        local = np.empty(self._local_shape, dtype=np.float32)
        self._check(self._lib.slmgs_comp_get_phase(self._ctx, _lib.fptr(local)))
        return self._comm.allgather_rows(local, self._rows, self._on_device, self._device)

    @phase.setter
    def phase(self, value):
        self.reset_phase(value)

    def reset_phase(self, custom_phase=None, random_phase=None, quadratic_phase=None):
        """_hologram.py:536-601; ``custom_phase`` has the shape of the whole SLM (or of this rank's slab)."""
        local = getattr(self, "_local_shape", None) or tuple(self.slm_shape)
        if custom_phase is not None:
            p = np.array(custom_phase, dtype=self.dtype)
            if tuple(p.shape) == tuple(getattr(self, "_full_shape", ())) and tuple(p.shape) != tuple(local):
                p = p[self._r0:self._r1]
            if tuple(p.shape) != tuple(local):
                raise ValueError(f"Reset phase of shape {p.shape} is not of slm_shape {self._full_shape}")
        else:
            if quadratic_phase or self.flags.get("quadratic_phase", False):
                raise NotImplementedError("quadratic_phase preconditioning is outside the GS/WGS hot path; pass phase=")
            if random_phase is None:
                random_phase = self.flags.get("random_phase", 1)
            p = np.zeros(local, dtype=self.dtype)
            if random_phase:
                p += random_phase * np.random.default_rng().uniform(-np.pi, np.pi, local).astype(self.dtype)
        self._check(self._lib.slmgs_comp_set_phase(self._ctx, _lib.fptr(_lib.f32(np.ascontiguousarray(p)))))
        self._phase_set = True
        self._amp_ff_set = False

    @property
    def nearfield(self):
        raise NotImplementedError("the near field of a sharded hologram lives in slabs; use .phase")

    def _reduce(self):
        if not self._on_device:
            self._check(self._lib.slmgs_comp_sync(self._ctx))
        self._comm.allreduce_f64(self._lib.slmgs_comp_facc_ptr(self._ctx), 2 * len(self), self._on_device, self._device,
                                 self._lib.slmgs_comp_stream(self._ctx))

    def _forward(self, populate):
        self._check(self._lib.slmgs_comp_near2far(self._ctx))
        self._reduce()
        self._check(self._lib.slmgs_comp_finalize(self._ctx, 1 if populate else 0))

    @property
    def farfield(self):
        if not self._amp_ff_set:
            self._forward(False)
            self._amp_ff_set = True
        return self._vec(self._lib.slmgs_comp_get_farfield, np.complex64)

    def optimize_gs(self, iterations, callback):
        """_hologram.py:1427-1493 with the compressed maps, the pixel sum completed by one all-reduce per iteration."""
        self._check_spot_zernike_change()
        mraf = self._mraf_enabled()
        if "WGS" in self.flags["method"]:
            self._check_feedback()
        for _ in iterations:
            if callback is not None:
                self._forward(False)
                self._amp_ff_set = True
                if callback(self):
                    break
            self._update_stats(self.flags["stat_groups"])
            params = self._iteration_params(mraf, stepped=callback is not None)
            self._check(self._lib.slmgs_comp_near2far(self._ctx))
            self._reduce()
            self._check(self._lib.slmgs_comp_constrain_far2near(self._ctx, C.byref(params)))
            self.iter += 1
        self._forward(True)
        self._amp_ff_set = True
        self._phase_ff_none = False
