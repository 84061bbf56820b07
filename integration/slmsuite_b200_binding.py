"""
Reference-side binding of ``libslmgs.so`` -- the file a slmsuite maintainer would drop into the reference tree as
``slmsuite/holography/algorithms/_b200.py``.

It needs nothing of this repository except the shared library and ``include/slmgs.h``'s contract: plain ctypes, and
one subclass of the reference's own ``Hologram`` that overrides ``optimize_gs`` (_hologram.py:1427-1493), the level at
which the reference's own subclasses extend the loop.  Everything else -- constructor, ``optimize()``, flags, stats,
``get_phase()`` -- is the unmodified reference code on its NumPy backend.

    B200Hologram = bind(reference_Hologram_class, "/path/to/libslmgs.so")

``tests/test_integration_reference.py`` executes exactly this file against the live reference.
"""
import ctypes as C

import numpy as np

_METHODS = {"GS": 0, "WGS-Leonardo": 1, "WGS-Kim": 2, "WGS-Nogrette": 3, "WGS-Wu": 4, "WGS-tanh": 5}
_PHASE_COMPUTE, _PHASE_COMPUTE_STORE, _PHASE_STORED = 0, 1, 2  # slmgs_phase_mode, include/slmgs.h


class _Params(C.Structure):  # slmgs_params, include/slmgs.h
    _fields_ = [("method", C.c_int), ("update_weights", C.c_int), ("phase_mode", C.c_int),
                ("feedback_exponent", C.c_float), ("feedback_factor", C.c_float),
                ("mraf", C.c_int), ("mraf_has_factor", C.c_int), ("mraf_factor", C.c_float),
                ("feedback", C.c_int), ("spot_width", C.c_int), ("zero_weights", C.c_int), ("zero_factor", C.c_float)]


def _load(path):
    lib = C.CDLL(path)
    f = C.POINTER(C.c_float)
    lib.slmgs_create.argtypes = [C.POINTER(C.c_void_p)] + [C.c_int] * 6
    lib.slmgs_destroy.argtypes = [C.c_void_p]
    lib.slmgs_last_error.restype = C.c_char_p
    lib.slmgs_last_error.argtypes = [C.c_void_p]
    for name in ("set_phase", "get_phase", "get_amp_ff", "get_phase_ff", "get_weights", "set_weights", "set_phase_ff",
                 "set_propagation"):
        getattr(lib, "slmgs_" + name).argtypes = [C.c_void_p, f]
    lib.slmgs_set_target.argtypes = [C.c_void_p, f, C.c_int]
    lib.slmgs_set_amp_array.argtypes = [C.c_void_p, f, C.c_int]
    lib.slmgs_set_amp_scalar.argtypes = [C.c_void_p, C.c_float]
    lib.slmgs_get_farfield.argtypes = [C.c_void_p, C.c_void_p]
    lib.slmgs_run.argtypes = [C.c_void_p, C.POINTER(_Params), C.c_int, C.c_int]
    return lib


def bind(reference_hologram, library_path):
    """Returns a subclass of the reference's ``Hologram`` whose GS / WGS loop runs in ``libslmgs.so``."""
    lib = _load(library_path)
    f = C.POINTER(C.c_float)

    def chk(ctx, status):
        if status:
            msg = lib.slmgs_last_error(ctx).decode()
            raise (ValueError if status == -1 else RuntimeError)(msg)

    def f32(a):
        return np.ascontiguousarray(a, dtype=np.float32)

    def ptr(a):
        return a.ctypes.data_as(f)

    class B200Hologram(reference_hologram):
        """The reference's Hologram with ``optimize_gs`` bound to libslmgs.so (NumPy backend for everything else)."""

        def _b200_iteration_params(self, mraf_enabled):
            """Host bookkeeping of ``_gs_farfield_routines`` (_hologram.py:1550-1605): whether the weights are updated
            this iteration, the WGS-Kim fixed-phase state machine on ``flags`` / ``stats["flags"]``, and from where
            the device takes the far-field phase."""
            fl = self.flags
            method = fl["method"]
            update = ("WGS" in method) and self.iter > 0                                        # :1552
            store = False
            if update:
                if "Kim" in method:                                                             # :1556-1583
                    was_not_fixed = not fl["fixed_phase"]
                    if fl["fix_phase_efficiency"] is not None:
                        raise ValueError("Must track statistics to fix phase based on efficiency!")
                    if was_not_fixed and self.iter >= fl["fix_phase_iteration"] - 1:
                        previous = self.stats["flags"]["fixed_phase"]
                        if all(not previous[-1 - i] for i in range(fl["fix_phase_iteration"])):
                            fl["fixed_phase"] = True
                    if (fl["fixed_phase"] and self._b200_phase_ff_none) or was_not_fixed:
                        store = True
                else:
                    fl["fixed_phase"] = False                                                   # :1584-1585
            if not fl.get("fixed_phase", False):
                mode = _PHASE_COMPUTE
            elif self._b200_phase_ff_none or store:
                mode = _PHASE_COMPUTE_STORE
            else:
                mode = _PHASE_STORED
            self._b200_phase_ff_none = False
            mf = fl.get("mraf_factor", None)
            return _Params(method=_METHODS[method], update_weights=int(update), phase_mode=mode,
                           feedback_exponent=float(fl.get("feedback_exponent", 0.0) or 0.0),
                           feedback_factor=float(fl.get("feedback_factor", 0.0) or 0.0),
                           mraf=int(mraf_enabled), mraf_has_factor=int(mf is not None),
                           mraf_factor=float(mf if mf is not None else 1.0),
                           feedback=0, spot_width=0, zero_weights=0, zero_factor=1.0)

        def optimize_gs(self, iterations, callback):
            fl = self.flags
            if (callback is not None or fl["stat_groups"] or fl.get("raw_stats", False)
                    or fl.get("feedback", "computational") != "computational"
                    or ("Kim" in fl["method"] and fl.get("fix_phase_efficiency", None) is not None)):
                return super().optimize_gs(iterations, callback)  # host code needs the far field between the transforms
            H, W = self.shape
            h, w = self.slm_shape
            ctx = C.c_void_p()
            chk(None, lib.slmgs_create(C.byref(ctx), 0, 1, int(H), int(W), int(h), int(w)))
            try:
                if np.isscalar(self.amp) or np.ndim(self.amp) == 0:
                    chk(ctx, lib.slmgs_set_amp_scalar(ctx, float(self.amp)))
                else:
                    chk(ctx, lib.slmgs_set_amp_array(ctx, ptr(f32(self.amp)), 0))
                if getattr(self, "propagation_kernel", None) is not None:
                    chk(ctx, lib.slmgs_set_propagation(ctx, ptr(f32(self.propagation_kernel))))
                chk(ctx, lib.slmgs_set_target(ctx, ptr(f32(self.target)), 0))
                chk(ctx, lib.slmgs_set_weights(ctx, ptr(f32(self.weights))))
                chk(ctx, lib.slmgs_set_phase(ctx, ptr(f32(self.phase))))
                self._b200_phase_ff_none = self.phase_ff is None
                if self.phase_ff is not None:
                    chk(ctx, lib.slmgs_set_phase_ff(ctx, ptr(f32(self.phase_ff))))
                mraf_enabled = bool(np.isnan(np.sum(self.target)))                              # :1495-1501
                plist = []
                for _ in iterations:                                                            # :1465-1490, host part
                    self._update_stats(fl["stat_groups"])                                       # :1479
                    plist.append(self._b200_iteration_params(mraf_enabled))                     # :1552-1585
                    self.iter += 1
                arr = (_Params * max(len(plist), 1))(*plist)
                chk(ctx, lib.slmgs_run(ctx, arr, len(plist), 1))                                # loop + _populate_results
                for name, attr, shape in (("get_phase", "phase", self.slm_shape), ("get_amp_ff", "amp_ff", self.shape),
                                          ("get_phase_ff", "phase_ff", self.shape), ("get_weights", "weights", self.shape)):
                    out = np.empty(shape, dtype=np.float32)
                    chk(ctx, getattr(lib, "slmgs_" + name)(ctx, ptr(out)))
                    cur = getattr(self, attr, None)
                    if isinstance(cur, np.ndarray) and cur.shape == out.shape and cur.dtype == out.dtype:
                        cur[...] = out                                                          # "modified in place", :70-77
                    else:
                        setattr(self, attr, out)
                ff = np.empty(self.shape, dtype=np.complex64)
                chk(ctx, lib.slmgs_get_farfield(ctx, ff.ctypes.data_as(C.c_void_p)))
                self.farfield[...] = ff
            finally:
                lib.slmgs_destroy(ctx)

    B200Hologram.__name__ = "B200Hologram"
    return B200Hologram
