"""
CPU oracle for the GS / WGS hologram loop  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the algorithm that the reference
(slmsuite v0.4.1 @ 39243f08) runs inside ``Hologram.optimize`` /
``SpotHologram.optimize`` on its NumPy backend.  It exists to *check* the CUDA
path; it is never imported by the product package ``slmsuite_b200``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(imported from /root/reference, this container only) side by side with this
restatement and commits the reference's outputs under ``tests/golden/``;
``tests/test_oracle.py`` replays them.  On NumPy 2.3.x the restatement
reproduces the reference bit for bit (same ufunc calls in the same order and
dtypes), see DESIGN.md "Oracle".

Every function cites the reference lines (relative to /root/reference/) it
follows.  The arithmetic of the FFT itself lives in NumPy's bundled pocketfft
(numpy 2.3.5 here; the reference pins no version), which both sides call.
"""

import numpy as np

# slmsuite/holography/algorithms/_header.py:53-81
METHOD_DEFAULTS = {
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
FEEDBACKS = ("computational", "computational_spot", "experimental", "experimental_spot", "external_spot")


# --------------------------------------------------------------------------- helpers
def l2norm(a):
    """sqrt(nansum(|a|^2)) in a's dtype.  _hologram.py:1980-2011."""
    if np.iscomplexobj(a):
        return np.sqrt(np.nansum(np.square(np.abs(a))))
    return np.sqrt(np.nansum(np.square(a)))


def crop_bounds(shape, slm_shape):
    """Centred crop (i0, i1, i2, i3).  toolbox/__init__.py:1665-1712."""
    dy = (shape[0] - slm_shape[0]) / 2.0
    dx = (shape[1] - slm_shape[1]) / 2.0
    if dy < 0 or dx < 0:
        raise ValueError(f"Shape {tuple(shape)} is too small to unpad to shape {tuple(slm_shape)}")
    return (
        int(np.floor(dy)),
        int(shape[0] - np.ceil(dy)),
        int(np.floor(dx)),
        int(shape[1] - np.ceil(dx)),
    )


def padded_shape(slm_shape, padding_order=1, square_padding=True):
    """Next power-of-two padding.  _hologram.py:712-725 (precision=inf branch)."""
    if padding_order > 0:
        shp = np.power(2, np.ceil(np.log2(slm_shape)) + padding_order - 1).astype(int)
    else:
        shp = np.asarray(slm_shape)
    shp = tuple(int(s) for s in shp)
    if square_padding:
        m = max(shp)
        shp = (m, m)
    return shp


def take_sum(image, centres_xy, width):
    """
    Window-integrated power round each spot: floor'ed integer centres, offsets
    floor(arange(w) - (w-1)/2), float64 accumulation, NumPy wrap-around for negative
    indices and IndexError past the end.  analysis/__init__.py:61-204 with
    centered=True, integrate=True, clip=False.
    """
    c = np.floor(np.asarray(centres_xy)).astype(int)
    off = np.floor(np.arange(width).astype(np.float64) - float(width - 1) / 2).astype(int)
    ox, oy = np.meshgrid(off, off)
    ix = ox.ravel()[np.newaxis, :] + c[0][:, np.newaxis]
    iy = oy.ravel()[np.newaxis, :] + c[1][:, np.newaxis]
    win = image[np.newaxis, iy, ix]
    return np.squeeze(np.sum(win.astype(float), axis=-1))


def disc_indices(cx, cy, w, shape):
    """
    Pixels of the circular window of diameter ``w`` centred on (cx, cy), clipped to
    ``shape``: toolbox/__init__.py:499-533 (window_slice, centered=True, circular=True) as
    called from _spots.py:1531-1538.
    """
    xi = int(cx - (w - 2) / 2)
    xf = xi + int(w)
    yi = int(cy - (w - 2) / 2)
    yf = yi + int(w)
    xi, xf = np.clip([xi, xf], 0, shape[1] - 1)
    yi, yf = np.clip([yi, yf], 0, shape[0] - 1)
    xg, yg = np.meshgrid(np.arange(xi, xf), np.arange(yi, yf))
    xc = xi + int((w - 1) / 2)
    yc = yi + int((w - 1) / 2)
    rr = (w ** 2) * np.square(xg.astype(float) - xc) + (w ** 2) * np.square(yg.astype(float) - yc)
    m = rr <= (w ** 2) * (w ** 2) / 4.0
    ys = np.clip(np.ravel(yg[m]), 0, shape[0] - 1)
    xs = np.clip(np.ravel(xg[m]), 0, shape[1] - 1)
    return ys, xs


def smallest_chebyshev(vectors):
    """Smallest pairwise inf-norm distance (toolbox/__init__.py:1127-1230), O(N log N)."""
    v = np.asarray(vectors, dtype=float)
    n = v.shape[1]
    if n < 2:
        return np.inf
    from scipy.spatial import cKDTree

    d, _ = cKDTree(v.T).query(v.T, k=2, p=np.inf)
    return float(np.min(d[:, 1]))


def calc_stats(feedback_amp, target_amp, total=None):
    """
    efficiency / uniformity / pkpk_err / std_err.  _stats.py:7-116 with
    efficiency_compensation=False, raw=False.  NOTE the reference normalises
    ``feedback_amp`` and ``target_amp`` IN PLACE (_stats.py:51-69); so does this.
    """
    fpw = np.square(feedback_amp)
    tpw = np.square(target_amp)
    if total is not None:
        eff = np.nansum(fpw) / total
    fsum = np.sum(fpw)
    fpw *= 1 / fsum
    feedback_amp *= 1 / np.sqrt(fsum)
    tsum = np.nansum(tpw)
    tpw *= 1 / tsum
    target_amp *= 1 / np.sqrt(tsum)
    if total is None:
        eff = np.square(float(np.nansum(np.multiply(target_amp, feedback_amp))))
    mask = np.logical_and(tpw != 0, np.logical_not(np.isnan(tpw)))
    fm = fpw[mask]
    tm = tpw[mask]
    ratio = np.divide(fm, tm)
    err = tm - fm
    rmin = float(np.amin(ratio))
    rmax = float(np.amax(ratio))
    return {
        "efficiency": float(eff),
        "uniformity": float(1 - (rmax - rmin) / (rmax + rmin)),
        "pkpk_err": float(err.size * float(np.amax(err) - np.amin(err))),
        "std_err": float(err.size * float(np.std(err))),
    }


def weight_multiplier_update(weights, feedback, target, method, flags, dtype):
    """
    In-place WGS update of ``weights`` and L2 renormalisation.
    _hologram.py:1822-1879 (``_update_weights_generic_cupy``, nan_checks=True).
    Returns the (modified) ``weights``.
    """
    m = method.lower()
    if m[:4] != "wgs-":
        raise ValueError("Weighting is only for WGS.")
    m = m[4:]

    fc = np.array(feedback, copy=True, dtype=dtype)
    fc *= 1 / l2norm(fc)

    if "wu" in m or "tanh" in m:  # additive family, :1833-1835
        fc *= -flags["feedback_exponent"]
        fc += np.asarray(target)
    else:  # multiplicative family, :1836-1843
        np.divide(fc, np.asarray(target), out=fc)
        fc[fc == np.inf] = 1
        fc[np.asarray(target) == 0] = 1
        np.nan_to_num(fc, copy=False, nan=1)

    if "leonardo" in m or "kim" in m:  # :1846-1848
        np.power(fc, -flags["feedback_exponent"], out=fc)
    elif "nogrette" in m:  # :1849-1855
        fc *= -(1 / np.nanmean(fc))
        fc += 1
        fc *= -flags["feedback_factor"]
        fc += 1
        np.reciprocal(fc, out=fc)
    elif "wu" in m:  # :1856-1857
        fc = np.exp(flags["feedback_exponent"] * fc)
    elif "tanh" in m:  # :1858-1860
        fc = flags["feedback_factor"] * np.tanh(flags["feedback_exponent"] * fc)
        fc += 1
    else:
        raise ValueError(f"Method '{method}' not recognized")

    fc[fc == np.inf] = 1  # :1866-1867
    weights *= fc  # :1870
    np.nan_to_num(weights, copy=False, nan=0.0001)  # :1872-1873
    weights *= 1 / l2norm(weights)  # :1877
    return weights


# --------------------------------------------------------------------------- Hologram
class OracleHologram:
    """
    State + loop of the reference ``Hologram`` on NumPy.
    State: _hologram.py:196-478.  Loop: _hologram.py:1427-1661.
    """

    def __init__(self, target, amp=None, phase=None, slm_shape=None, dtype=np.float32,
                 propagation_kernel=None, **flags):
        # shape voting, _hologram.py:296-356 (array / tuple inputs only)
        cands = []
        for a in (amp, phase):
            if a is not None:
                cands.append(tuple(np.shape(a)))
        if slm_shape is not None:
            cands.append(tuple(int(s) for s in slm_shape))
        if cands and any(c != cands[0] for c in cands):
            raise ValueError("amp / phase / slm_shape shapes disagree")
        self.slm_shape = cands[0] if cands else None

        # target / shape, _hologram.py:358-387
        if len(target) == 2 and np.ndim(target) == 1:
            self.shape = (int(target[0]), int(target[1]))
            target = None
        elif np.ndim(target) == 2:
            self.shape = tuple(np.shape(target))
        else:
            raise ValueError(f"Unexpected target {target}.")
        if self.slm_shape is None:
            self.slm_shape = self.shape

        # dtype, _hologram.py:391-398
        if dtype(0).nbytes == 4:
            self.dtype, self.dtype_complex = np.float32, np.complex64
        elif dtype(0).nbytes == 8:
            self.dtype, self.dtype_complex = np.float64, np.complex128
        else:
            raise ValueError(f"Data type {dtype} not supported.")

        # amp, _hologram.py:401-405  (scalar amp is an np.float64!)
        if amp is None:
            self.amp = 1 / np.sqrt(np.prod(self.slm_shape))
        else:
            # copy=None: no copy when the caller's array is already of this dtype, so the normalisation below
            # happens IN the caller's array (reference quirk, visible when children share one amp array)
            self.amp = np.array(amp, dtype=self.dtype, copy=None)
            self.amp *= 1 / l2norm(self.amp)

        # propagation kernel, _hologram.py:408-415
        if propagation_kernel is None:
            self.propagation_kernel = None
        else:
            self.propagation_kernel = np.array(propagation_kernel, dtype=self.dtype)
            if self.propagation_kernel.shape != self.slm_shape:
                raise ValueError("Expected the propagation kernel to be the same shape as the SLM.")

        self.flags = dict(flags)
        self._set_target(target)

        # phase, _hologram.py:570-601 (explicit phase, or a seeded uniform draw)
        self.phase = np.zeros(self.slm_shape, dtype=self.dtype)
        self.reset_phase(phase)
        self.reset(reset_phase=False)

    # -- construction helpers
    def _set_target(self, new_target):
        """_hologram.py:760-766."""
        if new_target is None:
            self.target = np.zeros(shape=self.shape, dtype=self.dtype)
        else:
            self.target = np.array(new_target, dtype=self.dtype)
            np.abs(self.target, out=self.target)
            with np.errstate(all="ignore"):
                self.target *= 1 / l2norm(self.target)

    def reset_phase(self, custom_phase=None, seed=None):
        """_hologram.py:570-601; the random branch takes an explicit seed here."""
        if custom_phase is not None:
            custom_phase = np.array(custom_phase, dtype=self.dtype)
            if tuple(custom_phase.shape) != tuple(self.slm_shape):
                raise ValueError(
                    f"Reset phase of shape {custom_phase.shape} is not of slm_shape {self.slm_shape}")
            np.copyto(self.phase, custom_phase)
        else:
            rng = np.random.default_rng(seed)
            self.phase[...] = rng.uniform(-np.pi, np.pi, self.slm_shape).astype(self.dtype)

    def reset_weights(self):
        """_hologram.py:603-614."""
        self.weights = self.target.copy()
        if hasattr(self, "zero_weights"):
            self.zero_weights *= 0
        np.nan_to_num(self.weights, copy=False, nan=0)

    def reset(self, reset_phase=True):
        """_hologram.py:442-478."""
        if reset_phase:
            self.reset_phase()
        self.reset_weights()
        self.iter = 0
        self.stats = {"method": [], "flags": {}, "stats": {}}
        self.amp_ff = None
        self.phase_ff = None
        self.nearfield = np.zeros(self.shape, dtype=self.dtype_complex)
        self.farfield = np.zeros(self.target.shape, dtype=self.dtype_complex)

    # -- accessors
    def get_phase(self):
        """_hologram.py:786-811 (no propagation)."""
        return self.phase + np.pi

    # -- transforms
    def _forward(self):
        """_hologram.py:1000-1011 + 1038-1056 + 951-953."""
        i0, i1, i2, i3 = crop_bounds(self.shape, self.slm_shape)
        self.nearfield.fill(0)
        if self.propagation_kernel is None:
            self.nearfield[i0:i1, i2:i3] = self.amp * np.exp(1j * self.phase)
        else:
            self.nearfield[i0:i1, i2:i3] = self.amp * np.exp(1j * (self.phase + self.propagation_kernel))
        self.farfield = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(self.nearfield), norm="ortho"))
        self.amp_ff = np.abs(self.farfield, out=self.amp_ff)

    def _inverse(self, extract=True):
        """_hologram.py:1058-1073 + 1026-1036."""
        i0, i1, i2, i3 = crop_bounds(self.shape, self.slm_shape)
        self.nearfield = np.fft.ifftshift(np.fft.ifft2(np.fft.ifftshift(self.farfield), norm="ortho"))
        if not extract:
            return
        self.phase = np.arctan2(self.nearfield.imag[i0:i1, i2:i3], self.nearfield.real[i0:i1, i2:i3],
                                out=self.phase)
        if self.propagation_kernel is not None:
            self.phase -= self.propagation_kernel

    def _populate_results(self):
        """_hologram.py:934-949."""
        self._forward()
        self.amp_ff = np.abs(self.farfield, out=self.amp_ff)
        self.phase_ff = np.arctan2(self.farfield.imag, self.farfield.real, out=self.phase_ff)

    # -- bookkeeping
    def _merge_flags(self, method, feedback, stat_groups, kw):
        """_hologram.py:1370-1410."""
        if method not in METHOD_DEFAULTS:
            raise ValueError("Unrecognized method '{}'.".format(method))
        self.flags["method"] = method
        for k, v in METHOD_DEFAULTS[method].items():
            if k not in self.flags:
                self.flags[k] = v
        if "fixed_phase" not in self.flags:
            self.flags["fixed_phase"] = False
        for k in kw:
            self.flags[k] = kw[k]
        for g in stat_groups:
            if g not in FEEDBACKS:
                raise ValueError("Statistics group '{}' not recognized as a feedback option.".format(g))
        self.flags["stat_groups"] = stat_groups
        if feedback is not None:
            if feedback not in FEEDBACKS:
                raise ValueError("Feedback '{}' not recognized as a feedback option.".format(feedback))
            self.flags["feedback"] = feedback

    def _stat_groups(self, groups):
        """_stats.py:118-128."""
        out = {}
        if "computational" in groups:
            out["computational"] = calc_stats(self.amp_ff, self.target)
        return out

    def _record(self, stats):
        """History lists, NaN padded.  _stats.py:130-190."""
        n = self.iter + 1
        meth = self.stats["method"]
        if len(meth) < n:
            meth.extend([""] * (n - len(meth)))
        meth[self.iter] = self.flags["method"]
        m = len(meth)
        fl = self.stats["flags"]
        for key in set(self.flags) | set(fl):
            if key not in fl:
                fl[key] = [np.nan] * m
            elif len(fl[key]) < n:
                fl[key].extend([np.nan] * (n - len(fl[key])))
            if key in self.flags:
                fl[key][self.iter] = self.flags[key]
        st = self.stats["stats"]
        groups = set(stats) | set(st)
        if groups:
            names = set()
            for g in stats:
                names |= set(stats[g])
            if st:
                names |= set(st[next(iter(st))])
            for g in groups:
                st.setdefault(g, {})
                for s in names:
                    if s not in st[g]:
                        st[g][s] = [np.nan] * m
                    elif len(st[g][s]) < n:
                        st[g][s].extend([np.nan] * (n - len(st[g][s])))
                    if g in stats and s in stats[g]:
                        st[g][s][self.iter] = stats[g][s]

    def _update_weights(self):
        """_hologram.py:1914-1922."""
        if self.flags["feedback"] == "computational":
            weight_multiplier_update(self.weights, self.amp_ff, self.target,
                                     self.flags["method"], self.flags, self.dtype)

    def _mraf_setup(self):
        """_hologram.py:1495-1548."""
        if not np.isnan(np.sum(self.target)):
            return None
        noise = np.isnan(self.target)
        zero = np.abs(self.target) == 0
        # complex accumulator over the zero region, created once and then always used (:1511-1515)
        if "zero_factor" in self.flags and self.flags["zero_factor"] != 0:
            if int(np.sum(zero)) > 0 and not hasattr(self, "zero_weights"):
                self.zero_weights = np.zeros((int(np.sum(zero)),), dtype=self.dtype_complex)
        signal = np.logical_not(np.logical_or(noise, zero))
        return {"noise": noise, "zero": zero, "signal": signal}

    def _constrain(self, mraf):
        """_hologram.py:1550-1653."""
        fl = self.flags
        if "WGS" in fl["method"] and self.iter > 0:
            self._update_weights()
            if "Kim" in fl["method"]:
                was_free = not fl["fixed_phase"]
                if fl["fix_phase_efficiency"] is not None:
                    st = self.stats["stats"]
                    if len(st) == 0:
                        raise ValueError("Must track statistics to fix phase based on efficiency!")
                    last = tuple(st.keys())[-1]
                    if st[last]["efficiency"][self.iter] > fl["fix_phase_efficiency"]:
                        fl["fixed_phase"] = True
                if was_free and self.iter >= fl["fix_phase_iteration"] - 1:
                    hist = self.stats["flags"]["fixed_phase"]
                    if all(not hist[-1 - i] for i in range(fl["fix_phase_iteration"])):
                        fl["fixed_phase"] = True
                if (fl["fixed_phase"] and self.phase_ff is None) or was_free:
                    self.phase_ff = np.arctan2(self.farfield.imag, self.farfield.real, out=self.phase_ff)
            else:
                fl["fixed_phase"] = False

        if mraf is None:
            if not fl.get("fixed_phase", False) or self.phase_ff is None:
                self.phase_ff = np.arctan2(self.farfield.imag, self.farfield.real, out=self.phase_ff)
            np.exp(1j * self.phase_ff, out=self.farfield)
            np.multiply(self.farfield, self.weights, out=self.farfield)
        else:
            if hasattr(self, "zero_weights"):  # :1613-1616
                fz = self.farfield[mraf["zero"]]
                self.zero_weights -= fl.get("zero_factor", 1) * np.abs(fz) * fz
                self.farfield[mraf["zero"]] = self.zero_weights
            else:
                self.farfield[mraf["zero"]] = 0
            if not fl.get("fixed_phase", False):
                self.phase_ff = np.arctan2(self.farfield.imag, self.farfield.real, out=self.phase_ff)
            np.exp(1j * self.phase_ff, where=mraf["signal"], out=self.farfield)
            np.multiply(self.farfield, self.weights, where=mraf["signal"], out=self.farfield)
            mf = fl.get("mraf_factor", None)
            if mf is not None:
                np.multiply(self.farfield, mf, where=mraf["noise"], out=self.farfield)

    # -- the loop
    def optimize(self, method="GS", maxiter=20, verbose=False, callback=None, feedback=None,
                 stat_groups=[], **kw):
        """_hologram.py:1351-1368 + 1427-1493."""
        kw.pop("name", None)
        self._merge_flags(method, feedback, stat_groups, kw)
        if "GS" not in method:
            raise ValueError(f"Unsupported optimization method '{method}'")
        mraf = self._mraf_setup()
        for _ in range(maxiter):
            self._forward()
            if callback is not None and callback(self):
                break
            self._record(self._stat_groups(self.flags["stat_groups"]))
            self._constrain(mraf)
            self._inverse()
            self.iter += 1
        self._populate_results()


# --------------------------------------------------------------------------- SpotHologram
class OracleSpotHologram(OracleHologram):
    """
    Reference ``SpotHologram`` with ``basis="knm"``, ``cameraslm=None``.
    ctor: _spots.py:1090-1373; targets: _spots.py:1490-1546; weights: :1573-1624.
    """

    def __init__(self, shape, spot_vectors, basis="knm", spot_amp=None, null_vectors=None,
                 null_radius=None, null_region=None, null_region_radius_frac=None, **kw):
        if basis not in (None, "knm"):
            raise ValueError("oracle supports basis='knm' only (other bases need a cameraslm)")
        v = np.squeeze(np.asarray(spot_vectors, dtype=float))
        if v.ndim == 1:
            v = v[:, np.newaxis]
        self.spot_knm = v
        n = v.shape[1]
        if spot_amp is not None:
            self.spot_amp = np.ravel(spot_amp)
            if len(self.spot_amp) != n:
                raise ValueError("spot_amp must have the same length as the provided spots.")
        else:
            self.spot_amp = np.full(n, 1.0 / np.sqrt(n))
        self.null_knm = None if null_vectors is None else np.asarray(null_vectors, dtype=float).reshape(2, -1)
        self.null_radius_knm = null_radius
        self.null_region_knm = null_region

        # integration width, _spots.py:1292-1297 (psf_knm = 0 without a cameraslm)
        dist = np.max([smallest_chebyshev(self.spot_knm) / 1.5, 3])
        width = np.clip(10 * 0, 3, dist)
        self.spot_integration_width_knm = int(2 * np.floor(width / 2) + 1)

        # bounds, _spots.py:1309-1323
        if (np.any(v[0] < 0) or np.any(v[1] < 0) or np.any(v[0] >= shape[1]) or np.any(v[1] >= shape[0])):
            raise ValueError("Spots outside SLM computational space bounds!")

        # null radius default, _spots.py:1341-1346
        if self.null_knm is not None:
            if self.null_radius_knm is None:
                self.null_radius_knm = smallest_chebyshev(np.hstack((self.null_knm, self.spot_knm))) / 4
            self.null_radius_knm = int(np.ceil(self.null_radius_knm))

        amp = kw.pop("amp", None)
        super().__init__(tuple(shape), amp=amp, **kw)

        # _spots.py:1360-1370
        if null_region_radius_frac is not None:
            if self.null_region_knm is None:
                self.null_region_knm = np.zeros(self.shape, dtype=bool)
            xl = np.linspace(-1, 1, self.null_region_knm.shape[0])
            yl = np.linspace(-1, 1, self.null_region_knm.shape[1])
            xg, yg = np.meshgrid(xl, yl)
            self.null_region_knm[np.square(xg) + np.square(yg) > null_region_radius_frac ** 2] = True

        self.set_target(reset_weights=True)

    @staticmethod
    def make_rectangular_array(shape, array_shape, array_pitch, array_center=None, basis="knm",
                               orientation_check=False, **kw):
        """_spots.py:1441-1488 (basis='knm')."""
        if np.isscalar(array_shape):
            array_shape = (int(array_shape), int(array_shape))
        if np.isscalar(array_pitch):
            array_pitch = (array_pitch, array_pitch)
        if array_center is None:
            array_center = (shape[1] / 2.0, shape[0] / 2.0)
        xe = (np.arange(array_shape[0]) - (array_shape[0] - 1) / 2.0) * array_pitch[0] + array_center[0]
        ye = (np.arange(array_shape[1]) - (array_shape[1] - 1) / 2.0) * array_pitch[1] + array_center[1]
        xg, yg = np.meshgrid(xe, ye, sparse=False, indexing="xy")
        xs, ys = xg.ravel(), yg.ravel()
        if orientation_check and len(xs) > 2:
            xs, ys = xs[:-2], ys[:-2]
        return OracleSpotHologram(shape, np.vstack((xs, ys)), basis=basis, **kw)

    def set_target(self, reset_weights=False):
        """_spots.py:1490-1546."""
        self.spot_knm_rounded = np.rint(self.spot_knm).astype(int)
        if self.null_knm is None:
            self.target.fill(0)
        else:
            self.target.fill(np.nan)
            if self.null_region_knm is not None:
                self.target[self.null_region_knm] = 0
            pts = np.hstack((self.null_knm, self.spot_knm))
            w = int(2 * self.null_radius_knm + 1)
            for i in range(pts.shape[1]):
                ys, xs = disc_indices(np.rint(pts[0, i]), np.rint(pts[1, i]), w, self.target.shape)
                self.target[ys, xs] = 0
        self.target[self.spot_knm_rounded[1, :], self.spot_knm_rounded[0, :]] = self.spot_amp
        self.target /= l2norm(self.target)
        if reset_weights:
            self.reset_weights()

    def _update_weights(self):
        """_spots.py:1573-1624."""
        fb = self.flags["feedback"]
        if fb == "computational":
            weight_multiplier_update(self.weights, self.amp_ff, self.target,
                                     self.flags["method"], self.flags, self.dtype)
            return
        if fb != "computational_spot":
            raise ValueError("Feedback '{}' not recognized.".format(fb))
        sy, sx = self.spot_knm_rounded[1, :], self.spot_knm_rounded[0, :]
        amp_fb = np.sqrt(take_sum(np.square(self.amp_ff), self.spot_knm_rounded,
                                  self.spot_integration_width_knm))
        self.weights[sy, sx] = weight_multiplier_update(
            self.weights[sy, sx], np.array(amp_fb, dtype=self.dtype), self.spot_amp,
            self.flags["method"], self.flags, self.dtype)

    def _stat_groups(self, groups):
        """_spots.py:1626-1697 (computational + computational_spot groups, NumPy branch)."""
        out = super()._stat_groups(groups)
        if "computational_spot" in groups:
            sy, sx = self.spot_knm_rounded[1, :], self.spot_knm_rounded[0, :]
            if tuple(self.shape) == tuple(self.slm_shape):
                out["computational_spot"] = calc_stats(
                    self.amp_ff[sy, sx], self.spot_amp, total=np.sum(np.square(self.amp_ff)))
            else:
                pw = np.square(self.amp_ff)
                out["computational_spot"] = calc_stats(
                    np.sqrt(take_sum(pw, self.spot_knm, self.spot_integration_width_knm)),
                    self.spot_amp, total=np.sum(pw))
        return out


# --------------------------------------------------------------------------- MultiplaneHologram
class OracleMultiplaneHologram:
    """
    Reference ``MultiplaneHologram`` (_multiplane.py:29-75, :174-180, :214-286): N child holograms that share one
    near-field phase; the parent sums the weighted complex child near fields and extracts one phase.
    """

    def __init__(self, holograms, weights=None):
        self.holograms = holograms
        first = holograms[0]
        self.slm_shape = tuple(first.slm_shape)
        self.shape = self.slm_shape
        self.dtype, self.dtype_complex = first.dtype, first.dtype_complex
        # parent state, _multiplane.py:62-75: amp and phase of the first child, shared by every child.  The parent
        # constructor normalises the first child's amplitude array AGAIN, in place (_hologram.py:404-405 with
        # copy=None), which can move its last bits; a scalar amp makes the reference constructor fail.
        self.amp = first.amp
        if not np.isscalar(self.amp):
            self.amp *= 1 / l2norm(self.amp)
        self.phase = np.array(first.phase, dtype=self.dtype)
        self.propagation_kernel = None
        self.target = None
        for h in holograms:
            h.amp = self.amp
            h.phase = self.phase
        if weights is None:
            weights = np.ones(len(holograms), dtype=self.dtype)
        self.weights = np.array(weights, dtype=self.dtype)
        self.weights /= l2norm(self.weights)
        self.flags = {}
        self.iter = 0
        self.stats = {"method": [], "flags": {}, "stats": {}}
        self.nearfield = np.zeros(self.slm_shape, dtype=self.dtype_complex)

    def __len__(self):
        return len(self.holograms)

    def get_phase(self):
        return self.phase + np.pi

    def _merge_flags(self, method, feedback, stat_groups, kw):
        """_multiplane.py:174-180: parent flags first, then pushed into every child."""
        OracleHologram._merge_flags(self, method, feedback, stat_groups, kw)
        for h in self.holograms:
            h.flags.update(self.flags)

    def _forward(self):
        """_multiplane.py:255-259."""
        for h in self.holograms:
            h.phase = self.phase
            h._forward()
            h.iter = self.iter

    def _inverse(self):
        """_multiplane.py:261-281."""
        self.nearfield.fill(0)
        for h, w in zip(self.holograms, self.weights):
            h._inverse(extract=False)
            i0, i1, i2, i3 = crop_bounds(h.shape, h.slm_shape)
            if h.propagation_kernel is None:
                self.nearfield += w * h.nearfield[i0:i1, i2:i3]
            else:
                self.nearfield += w * h.nearfield[i0:i1, i2:i3] * np.exp(-1j * h.propagation_kernel)
            h.iter = self.iter
        self.phase = np.arctan2(self.nearfield.imag, self.nearfield.real, out=self.phase)

    def optimize(self, method="GS", maxiter=20, verbose=False, callback=None, feedback=None,
                 stat_groups=[], **kw):
        """_hologram.py:1351-1368 + 1427-1493 with the overrides of _multiplane.py:232-286."""
        kw.pop("name", None)
        self._merge_flags(method, feedback, stat_groups, kw)
        if "GS" not in method:
            raise ValueError(f"Unsupported optimization method '{method}'")
        mraf = [h._mraf_setup() for h in self.holograms]
        for _ in range(maxiter):
            self._forward()
            if callback is not None and callback(self):
                break
            for h in self.holograms:
                h._record(h._stat_groups(self.flags["stat_groups"]))
            for h, m in zip(self.holograms, mraf):
                h._constrain(m)
            self._inverse()
            self.iter += 1
        self._forward()  # _populate_results: the children refresh their far fields
