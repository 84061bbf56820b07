"""
``SpotHologram``: host-side mirror of ``slmsuite.holography.algorithms.SpotHologram``
(_spots.py:1021-1697) for DFT-grid spot arrays (``basis="knm"``, ``cameraslm=None``), on top of
the GPU ``Hologram``.  Per-spot window feedback (``feedback="computational_spot"``) runs in
device kernels (``slmgs_update_weights_spot``); camera feedback and the ``"kxy"`` / ``"ij"`` bases
need hardware objects and are outside this path.
"""

import ctypes as C

import numpy as np

from . import _lib
from .hologram import Hologram, _norm, calculate_stats


def format_2vectors(vectors):
    """(2, N) float array from list / tuple input: toolbox.format_2vectors (toolbox/__init__.py)."""
    v = np.array(vectors, dtype=float)
    v = np.squeeze(v)
    if v.ndim == 1:
        v = v[:, np.newaxis]
    if v.ndim != 2 or v.shape[0] != 2:
        raise ValueError(f"Expected a (2, N) array of vectors, got shape {v.shape}")
    return v


def smallest_distance(vectors):
    """
    Smallest pairwise Chebyshev (inf-norm) distance between the columns of ``vectors``
    (toolbox.smallest_distance with its default metric, toolbox/__init__.py:1127-1230).  A k-d tree
    keeps this O(N log N) for the 10k-spot case instead of the reference's O(N^2) ``pdist``.
    """
    v = np.asarray(vectors, dtype=float)
    if v.shape[1] < 2:
        return np.inf
    from scipy.spatial import cKDTree

    d, _ = cKDTree(v.T).query(v.T, k=2, p=np.inf)
    return float(np.min(d[:, 1]))


def circular_window_indices(cx, cy, w, shape):
    """
    (y, x) pixel lists of the disc of diameter ``w`` centred on (cx, cy), clipped to ``shape``:
    toolbox.window_slice(window=(cx, w, cy, w), centered=True, circular=True)
    (toolbox/__init__.py:499-533), as used by toolbox.imprint from _spots.py:1531-1538.
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


class SpotHologram(Hologram):
    """
    Spot-array hologram on the DFT grid (reference ``SpotHologram``, _spots.py:1021-1697).

    ``SpotHologram(shape, spot_vectors, basis="knm", spot_amp=None, cameraslm=None,
    null_vectors=None, null_radius=None, null_region=None, null_region_radius_frac=None, **kwargs)``
    (_spots.py:1090-1102; the reference's default ``basis="kxy"`` needs a ``cameraslm``).
    """

    def __init__(self, shape, spot_vectors, basis="knm", spot_amp=None, cameraslm=None,
                 null_vectors=None, null_radius=None, null_region=None, null_region_radius_frac=None,
                 **kwargs):
        if cameraslm is not None:
            raise NotImplementedError("cameraslm-based SpotHologram (camera feedback) is outside the GS/WGS hot path")
        if basis is not None and basis != "knm":
            if basis in ("kxy", "ij"):
                raise AssertionError("We need a cameraslm to interpret {}.".format(basis))
            raise Exception("Unrecognized basis for spots '{}'.".format(basis))

        vectors = format_2vectors(spot_vectors)
        self.cameraslm = None
        if spot_amp is not None:
            self.spot_amp = np.ravel(spot_amp)
            if len(self.spot_amp) != vectors.shape[1]:
                raise ValueError("spot_amp must have the same length as the provided spots.")
        else:
            self.spot_amp = np.full(vectors.shape[1], 1.0 / np.sqrt(vectors.shape[1]))
        self.external_spot_amp = np.copy(self.spot_amp)

        # knm basis, _spots.py:1170-1197
        self.spot_knm = vectors
        self.spot_kxy = None
        self.spot_ij = None
        self.null_knm = None if null_vectors is None else format_2vectors(null_vectors)
        self.null_radius_knm = null_radius
        self.null_region_knm = null_region

        # integration width, _spots.py:1268-1297 (psf_knm = 0 without a cameraslm)
        min_psf = 3
        dist_knm = np.max([smallest_distance(self.spot_knm) / 1.5, min_psf])
        width = np.clip(10 * 0, min_psf, dist_knm)
        self.spot_integration_width_knm = int(2 * np.floor(width / 2) + 1)
        self.spot_integration_width_ij = None

        # bounds, _spots.py:1309-1323
        if (np.any(self.spot_knm[0] < 0) or np.any(self.spot_knm[1] < 0)
                or np.any(self.spot_knm[0] >= shape[1]) or np.any(self.spot_knm[1] >= shape[0])):
            raise ValueError(
                "Spots outside SLM computational space bounds!\nSpots:\n{}\nBounds: {}".format(self.spot_knm, shape)
            )

        # null radius, _spots.py:1341-1346
        if self.null_knm is not None:
            if self.null_radius_knm is None:
                self.null_radius_knm = smallest_distance(np.hstack((self.null_knm, self.spot_knm))) / 4
            self.null_radius_knm = int(np.ceil(self.null_radius_knm))

        # FeedbackHologram / Hologram constructor with a zero target, _spots.py:1349, _feedback.py:93-100
        super().__init__(target=tuple(shape), **kwargs)
        self.img_ij = None
        self.img_knm = None

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
                               orientation_check=False, **kwargs):
        """_spots.py:1387-1488 (``basis="knm"``)."""
        if np.isscalar(array_shape):
            array_shape = (int(array_shape), int(array_shape))
        if np.isscalar(array_pitch):
            array_pitch = (array_pitch, array_pitch)
        if array_center is None:
            if basis == "knm":
                array_center = (shape[1] / 2.0, shape[0] / 2.0)
            else:
                raise AssertionError("We need a cameraslm to interpret {}.".format(basis))
        x_edge = (np.arange(array_shape[0]) - (array_shape[0] - 1) / 2.0) * array_pitch[0] + array_center[0]
        y_edge = (np.arange(array_shape[1]) - (array_shape[1] - 1) / 2.0) * array_pitch[1] + array_center[1]
        x_grid, y_grid = np.meshgrid(x_edge, y_edge, sparse=False, indexing="xy")
        x_list, y_list = x_grid.ravel(), y_grid.ravel()
        if orientation_check and len(x_list) > 2:
            x_list = x_list[:-2]
            y_list = y_list[:-2]
        return SpotHologram(shape, np.vstack((x_list, y_list)), basis=basis, spot_amp=None, **kwargs)

    # ------------------------------------------------------------------ target
    def _set_target_spots(self, reset_weights=False):
        """_spots.py:1490-1546."""
        self.spot_knm_rounded = np.rint(self.spot_knm).astype(int)
        self.spot_kxy_rounded = None
        self.spot_ij_rounded = None
        t = self._target
        if self.null_knm is None:
            t.fill(0)
        else:
            t.fill(np.nan)
            if self.null_region_knm is not None:
                t[self.null_region_knm] = 0
            all_spots = np.hstack((self.null_knm, self.spot_knm))
            w = int(2 * self.null_radius_knm + 1)
            for ii in range(all_spots.shape[1]):
                ys, xs = circular_window_indices(np.rint(all_spots[0, ii]), np.rint(all_spots[1, ii]), w, t.shape)
                t[ys, xs] = 0
        t[self.spot_knm_rounded[1, :], self.spot_knm_rounded[0, :]] = self.spot_amp
        t /= _norm(t)
        self._upload_target()
        sx = np.ascontiguousarray(self.spot_knm_rounded[0, :], dtype=np.int32)
        sy = np.ascontiguousarray(self.spot_knm_rounded[1, :], dtype=np.int32)
        self._check(self._lib.slmgs_set_spots(self._ctx, len(sx), _lib.iptr(sx), _lib.iptr(sy),
                                              _lib.fptr(_lib.f32(self.spot_amp))))
        if reset_weights:
            self.reset_weights()

    def set_target(self, reset_weights=False, plot=False):
        """_spots.py:1548-1571."""
        self._set_target_spots(reset_weights=reset_weights)

    # ------------------------------------------------------------------ weights
    def _device_feedbacks(self):
        return ("computational", "computational_spot")

    def _feedback_params(self):
        if self.flags.get("feedback") == "computational_spot":
            return 1, int(self.spot_integration_width_knm)
        return 0, 0

    def _check_windows(self):
        """analysis.take(clip=False) semantics (analysis/__init__.py:133-183): negative window indices wrap,
        indices past the end raise IndexError.  Checked on the host before the device gathers."""
        w = int(self.spot_integration_width_knm)
        hi = (-((w - 1) // 2) if w % 2 else -(w // 2)) + w - 1
        x, y = self.spot_knm_rounded[0], self.spot_knm_rounded[1]
        if np.any(x + hi >= self.shape[1]) or np.any(y + hi >= self.shape[0]):
            bad = int(max(np.max(x + hi) - self.shape[1], np.max(y + hi) - self.shape[0])) + max(self.shape)
            raise IndexError(f"index {bad} is out of bounds for the spot integration windows of width {w} "
                             f"in a far field of shape {self.shape}")

    def _iteration_params(self, mraf, stepped):
        p = super()._iteration_params(mraf, stepped)
        if p.update_weights and p.feedback == 1:
            self._check_windows()
        return p

    def _update_weights(self, params):
        """_spots.py:1573-1624."""
        feedback = self.flags["feedback"]
        if feedback == "computational":
            self._check(self._lib.slmgs_update_weights(self._ctx, C.byref(params)))
        elif feedback == "computational_spot":
            self._check(self._lib.slmgs_update_weights_spot(self._ctx, C.byref(params),
                                                            int(self.spot_integration_width_knm)))
        elif feedback in ("experimental", "experimental_spot", "external_spot"):
            raise NotImplementedError(f"Feedback '{feedback}' needs camera hardware (outside the GS/WGS hot path)")
        else:
            raise ValueError("Feedback '{}' not recognized.".format(feedback))

    # ------------------------------------------------------------------ statistics
    def _window_power(self, centres_xy, width):
        c = np.floor(np.asarray(centres_xy)).astype(np.int32)  # analysis.take floors, analysis/__init__.py:133
        sx = np.ascontiguousarray(c[0])
        sy = np.ascontiguousarray(c[1])
        out = np.zeros(len(sx), dtype=np.float64)
        total = np.zeros(1, dtype=np.float64)
        self._check(self._lib.slmgs_window_power(self._ctx, len(sx), _lib.iptr(sx), _lib.iptr(sy), int(width),
                                                 _lib.dptr(out), _lib.dptr(total)))
        return out, float(total[0])

    def _calculate_stats_computational_spot(self, stats, stat_groups=[]):
        """_spots.py:1626-1679."""
        if "computational_spot" in stat_groups:
            if tuple(self.shape) == tuple(self.slm_shape):
                pw, total = self._window_power(self.spot_knm_rounded, 1)
            else:
                pw, total = self._window_power(self.spot_knm, self.spot_integration_width_knm)
            stats["computational_spot"] = calculate_stats(
                np.sqrt(pw).astype(self.dtype), np.array(self.spot_amp, dtype=float), total=total)

    def _update_stats(self, stat_groups=[]):
        """_spots.py:1681-1697."""
        stats = {}
        self._calculate_stats_computational(stats, stat_groups)
        self._calculate_stats_computational_spot(stats, stat_groups)
        self._update_stats_dictionary(stats)
