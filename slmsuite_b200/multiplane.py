"""
``MultiplaneHologram``: several child holograms (planes of focus, point sets) that share ONE near-field
phase, mirroring ``slmsuite.holography.algorithms.MultiplaneHologram`` (_multiplane.py:8-289).
SURVEY.md 8f rank 1: "batch + one reduction over the batch axis".

Per iteration every child transforms the shared near field to its own far field, applies its own
constraint / weight update, transforms back WITHOUT extracting a phase, and the parent phase is the
argument of the weighted complex sum of the child near fields (with each child's propagation kernel
removed), _multiplane.py:255-281.  On the device each child is its own context (any padded shape); the
children share one stream, the sum lives in one complex (h, w) buffer and is formed by the row kernel of
each child's inverse transform (``slmgs_constrain_accumulate``).
"""

import ctypes as C

import numpy as np

from . import _lib
from .hologram import Hologram, _norm


class MultiplaneHologram:
    """
    ``MultiplaneHologram(holograms, weights=None)`` (_multiplane.py:42-75).

    holograms : list of ``slmsuite_b200.Hologram`` / ``SpotHologram`` with identical ``slm_shape`` on one device
    weights   : N floats (power split between the children), L2-normalised; default equal
    """

    def __init__(self, holograms, weights=None):
        self.holograms = list(holograms)
        if len(self.holograms) == 0:
            raise ValueError("MultiplaneHologram needs at least one child hologram")
        for h in self.holograms:
            if isinstance(h, MultiplaneHologram):
                raise ValueError("Multiplane hologram recursion is not supported.")
            if not isinstance(h, Hologram):
                raise ValueError(f"Multiplane hologram must be provided child holograms, not {type(h)}")
        first = self.holograms[0]
        for h in self.holograms[1:]:
            if tuple(h.slm_shape) != tuple(first.slm_shape):
                raise ValueError("All child holograms must share one slm_shape")
            if h._device != first._device or h._batch_size() != first._batch_size():
                raise ValueError("All child holograms must live on the same device with the same batch size")
        self._lib = first._lib
        self.slm_shape = tuple(first.slm_shape)
        self.shape = self.slm_shape  # the parent has a fake target of slm_shape, _multiplane.py:63
        self.dtype, self.dtype_complex = first.dtype, first.dtype_complex
        self.target = None
        self.propagation_kernel = None

        # the children point to the parent's amp and phase (= the first child's), _multiplane.py:62-75
        self._amp = first.amp
        if not np.isscalar(self._amp):
            a = np.array(self._amp, dtype=self.dtype)
            a *= 1 / _norm(a)  # the reference parent normalises the first child's array once more
            self._amp = a
        phase0 = first.phase
        for h in self.holograms:
            h._check(self._lib.slmgs_share_stream(h._ctx, first._ctx))
            if np.isscalar(self._amp):
                h._check(self._lib.slmgs_set_amp_scalar(h._ctx, float(self._amp)))
            else:
                h._check(self._lib.slmgs_set_amp_array(h._ctx, _lib.fptr(_lib.f32(self._amp)), 0))
            h._amp = self._amp
            h.reset_phase(phase0)

        if weights is None:
            weights = np.ones(len(self), dtype=self.dtype)
        self.weights = np.array(weights, dtype=self.dtype)
        if self.weights.shape != (len(self),):
            raise ValueError("weights must hold one float per child hologram")
        self.weights /= _norm(self.weights)

        self._sum = self._lib.slmgs_nearfield_sum_ptr(first._ctx)
        if not self._sum:
            raise MemoryError("could not allocate the multiplane near-field accumulator")
        self.flags = {}
        self.iter = 0
        self.stats = {"method": [], "flags": {}, "stats": {}}

    def __len__(self):
        return len(self.holograms)

    # ------------------------------------------------------------------ shared state
    @property
    def amp(self):
        return self._amp

    @property
    def phase(self):
        return self.holograms[0].phase

    @phase.setter
    def phase(self, value):
        self.reset_phase(value)

    def get_phase(self, include_propagation=False):
        """_hologram.py:786-811 (the parent has no propagation kernel)."""
        return self.phase + np.pi

    extract_phase = get_phase

    def get_phase_gray(self, bitdepth=8, phase_correction=None):
        return self.holograms[0].get_phase_gray(bitdepth, phase_correction)

    def get_amp(self):
        return self._amp

    def reset_phase(self, custom_phase=None, random_phase=None, quadratic_phase=None):
        """The phase is shared: resetting the parent's resets every child's (_multiplane.py:214-220)."""
        first = self.holograms[0]
        first.reset_phase(custom_phase, random_phase, quadratic_phase)
        p = first.phase
        for h in self.holograms[1:]:
            h.reset_phase(p)

    def reset(self, reset_phase=True, reset_flags=False):
        """_multiplane.py:214-220."""
        if reset_phase:
            self.reset_phase()
        self.iter = 0
        self.stats = {"method": [], "flags": {}, "stats": {}}
        if reset_flags:
            self.flags = {"method": ""}
        for h in self.holograms:
            h.reset(reset_phase=False, reset_flags=reset_flags)

    def reset_weights(self):
        for h in self.holograms:
            h.reset_weights()

    def set_target(self, *args, **kwargs):
        """_multiplane.py:238-242."""
        raise RuntimeError(
            "Do not use MultiplaneHologram.set_target(). "
            "Instead, update the targets of the children holograms directly."
        )

    # ------------------------------------------------------------------ optimisation
    def _update_flags(self, method, verbose, feedback, stat_groups, **kwargs):
        """_multiplane.py:174-180: parent flags first, then pushed into every child."""
        Hologram._update_flags(self, method, verbose, feedback, stat_groups, **kwargs)
        for h in self.holograms:
            h.flags.update(self.flags)

    def _update_stats(self, stat_groups=[]):
        """_multiplane.py:232-234."""
        for h in self.holograms:
            h._update_stats(stat_groups)

    def optimize(self, method="GS", maxiter=20, verbose=True, callback=None, feedback=None,
                 stat_groups=[], **kwargs):
        """_hologram.py:1351-1368 with the overrides of _multiplane.py:255-286."""
        kwargs.pop("name", None)
        self._update_flags(method, verbose, feedback, stat_groups, **kwargs)
        iterations = range(maxiter)
        if verbose and maxiter > 1:
            try:
                from tqdm.auto import tqdm
                iterations = tqdm(iterations)
            except Exception:
                pass
        if "GS" not in method:
            raise ValueError(f"Unsupported optimization method '{method}'")
        self.optimize_gs(iterations, callback)

    def _forward_children(self):
        """_multiplane.py:255-259."""
        for h in self.holograms:
            h._check(self._lib.slmgs_forward(h._ctx))
            h._amp_ff_set = True
            h.iter = self.iter

    def optimize_gs(self, iterations, callback):
        """_hologram.py:1427-1493 with _multiplane.py:255-286."""
        mraf = [h._mraf_enabled() for h in self.holograms]
        if all(h._fusable(callback) for h in self.holograms):
            # nothing on the host looks at the far fields between the transforms: every child runs one fused
            # iteration (row first + fused column kernel) whose row inverse accumulates into the shared sum
            for _ in iterations:
                first = 1
                for h, w, m in zip(self.holograms, self.weights, mraf):
                    h.iter = self.iter
                    h._update_stats(self.flags["stat_groups"])
                    params = h._iteration_params(m, stepped=False)
                    h._check(self._lib.slmgs_run_accumulate(h._ctx, C.byref(params), float(w),
                                                            C.c_void_p(self._sum), first))
                    first = 0
                for h in self.holograms:
                    h._check(self._lib.slmgs_extract_phase_from_sum(h._ctx, C.c_void_p(self._sum)))
                self.iter += 1
            self._forward_children()
            return

        for _ in iterations:
            self._forward_children()  # (A)
            if callback is not None:  # (B.1)
                if callback(self):
                    break
            self._update_stats(self.flags["stat_groups"])  # (B.2)
            first = 1
            for h, w, m in zip(self.holograms, self.weights, mraf):  # (B.3) + (C): sum of the child near fields
                params = h._iteration_params(m, stepped=True)
                if params.update_weights:
                    h._update_weights(params)
                h._check(self._lib.slmgs_constrain_accumulate(h._ctx, C.byref(params), float(w),
                                                              C.c_void_p(self._sum), first))
                first = 0
                h.iter = self.iter
            for h in self.holograms:  # _nearfield_extract of the parent, shared by the children
                h._check(self._lib.slmgs_extract_phase_from_sum(h._ctx, C.c_void_p(self._sum)))
            self.iter += 1
        self._forward_children()  # _populate_results: the children refresh their far fields
