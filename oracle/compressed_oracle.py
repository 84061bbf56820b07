"""
CPU oracle for ``CompressedSpotHologram``  --  TEST INFRASTRUCTURE ONLY.

NumPy restatement of the reference's kernel-based ("compressed") spot hologram on its NumPy backend
(slmsuite/holography/algorithms/_spots.py:178-1019, option 1 of its docstring: cached kernels + matmul).
SURVEY.md 8f rank 4.  Instead of a DFT grid, every spot n owns a phase kernel

    kernel[n, pix] = exp(i * sum_d a[d, n] * Z_d(x_pix, y_pix)) / sqrt(H W)          (_spots.py:595-636)

with Z_d the Zernike polynomials of ``zernike_basis`` (ANSI indices) on the aperture-scaled SLM grid, and

    farfield = conj(kernel @ conj(nearfield)),  farfield /= ||farfield||             (_spots.py:767-824)
    nearfield = farfield @ kernel                                                    (_spots.py:887-915)

The GS / WGS loop around these two maps is the base ``Hologram``'s (``OracleHologram``), acting on N-vectors.

Zernike polynomials -> monomials: ``zernike_monomials`` restates toolbox/phase.py:1357-1420 (the combinatorial
expansion of doi:10.1117/12.294412) with the ANSI -> (n, l) map of :603-605.

Parity status: PINNED WITH A TOLERANCE.  ``oracle/make_golden_compressed.py`` runs the unmodified reference
(``CompressedSpotHologram`` on a ``SimulatedSLM`` / ``SimulatedCamera`` / ``FourierSLM`` with an explicit Fourier
calibration) and records its results; this restatement evaluates the kernel phase in float64 and rounds once to
float32, whereas the reference accumulates monomials in complex64 (toolbox/phase.py ``polynomial``), so the two differ
by float32 rounding of the phase: measured rel-RMSE(|farfield|) <= 3e-7, phase rms <= 7e-6 rad, weights <= 5e-7 on
the five golden cases (the generating script asserts 1e-5 / 1e-4 / 1e-5); Zernike coefficients identical.
"""
from math import factorial

import numpy as np

from oracle import gs_oracle


# --------------------------------------------------------------------------- Zernike -> monomials
def ansi_to_radial(index):
    """toolbox/phase.py:603-605."""
    n = int(np.floor(0.5 * np.sqrt(8 * index + 1) - 0.5))
    l = 2 * index - n * (n + 2)
    return n, l


def zernike_monomials(index):
    """{(a, b): coefficient} of x^a y^b for the real Zernike polynomial of ANSI ``index``;
    toolbox/phase.py:1357-1420."""
    n, l = ansi_to_radial(int(index))
    l = -l
    if l % 2:
        q = int((abs(l) - 1) / 2)
    elif l > 0:
        q = int(abs(l) / 2 - 1)
    else:
        q = int(abs(l) / 2)
    p = 0 if l <= 0 else 1
    l = abs(l)
    m = int((n - l) / 2)

    def comb(a, b):
        return factorial(a) / (factorial(b) * factorial(a - b))

    out = {}
    for i in range(q + 1):
        for j in range(m + 1):
            for k in range(m - j + 1):
                factor = -1 if (i + j) % 2 else 1
                factor *= comb(l, 2 * i + p)
                factor *= comb(m - j, k)
                factor *= float(factorial(n - j)) / (factorial(j) * factorial(m - j) * factorial(n - m - j))
                key = (int(n - 2 * (i + j + k) - p), int(2 * (i + k) + p))
                out[key] = out.get(key, 0) + int(factor)
    return {k: v for k, v in out.items() if v != 0}


def default_basis(D):
    """toolbox/phase.py:923-962 (``_zernike_indices_parse(None, D)``)."""
    if D == 2:
        return np.array([2, 1])
    if D == 3:
        return np.array([2, 1, 4])
    if D == 4:
        return np.array([2, 1, 4, 3])
    return np.hstack((np.array([2, 1, 4, 3]), np.arange(5, D + 1)))


def monomial_table(zernike_basis):
    """(px[M], py[M], c[M][D]): Z_d = sum_m c[m, d] x^px[m] y^py[m] for the basis (no negative / special indices)."""
    terms = {}
    for d, idx in enumerate(zernike_basis):
        if idx < 0:
            raise ValueError("special (negative) Zernike indices are not supported")
        for key, coef in zernike_monomials(idx).items():
            terms.setdefault(key, np.zeros(len(zernike_basis)))[d] = coef
    keys = sorted(terms, key=lambda ab: (ab[0] + ab[1], ab[1]))
    px = np.array([k[0] for k in keys], dtype=np.int32)
    py = np.array([k[1] for k in keys], dtype=np.int32)
    c = np.array([terms[k] for k in keys], dtype=np.float64).reshape(len(keys), len(zernike_basis))
    return px, py, c


def kxy_to_zernike(vectors, zernike_scaling):
    """toolbox.convert_vector(from "kxy", to "zernike"), toolbox/__init__.py:312-316, :355-356, :390-391."""
    v = np.array(vectors, dtype=float)
    scale = 2 * np.pi * np.reciprocal(zernike_scaling)
    out = v.copy()
    out[:2] = v[:2] * scale
    if v.shape[0] > 2:
        out[2] = v[2] * ((scale * scale) / (8 * np.pi))
    return out


# --------------------------------------------------------------------------- the hologram
class OracleCompressedSpotHologram(gs_oracle.OracleHologram):
    """
    ``CompressedSpotHologram(spot_vectors, basis, spot_amp, cameraslm)`` with the hardware object replaced by the two
    things read from it: ``slm_grid`` = ``cameraslm.slm.grid`` (x_grid, y_grid) and ``zernike_scaling`` =
    ``cameraslm.slm.get_source_zernike_scaling()`` (_spots.py:609-614 via ``zernike_aperture``).
    """

    def __init__(self, spot_vectors, basis="kxy", spot_amp=None, slm_grid=None, zernike_scaling=None,
                 amp=None, phase=None, **flags):
        spot_vectors = np.array(spot_vectors, dtype=float)
        D, N = spot_vectors.shape
        self.spot_amp = np.full(N, 1.0 / np.sqrt(N)) if spot_amp is None else np.array(spot_amp)
        if self.spot_amp.size != N:
            raise ValueError("spot_amp must have the same length as the provided spots")
        if isinstance(basis, str):
            self.zernike_basis = default_basis(D)
        else:
            self.zernike_basis = np.ravel(basis)
            basis = "zernike"
            if len(self.zernike_basis) != D:
                raise ValueError("zernike_basis must have the same dimension as the provided spots")
        if not np.any(self.zernike_basis == 2) or not np.any(self.zernike_basis == 1):
            raise ValueError("Compressed basis must include x, y (Zernike ANSI indices 2, 1)")
        if basis == "zernike":
            self.spot_zernike = spot_vectors
        elif basis == "kxy":
            self.spot_zernike = kxy_to_zernike(spot_vectors, zernike_scaling)
        else:
            raise ValueError("basis must be 'kxy', 'zernike' or a list of ANSI indices here")
        x_grid, y_grid = slm_grid
        slm_shape = x_grid.shape
        # _spots.py:609-614: grids pre-scaled by the aperture, stored as complex64 (i.e. rounded to float32)
        self._x = np.array(x_grid * zernike_scaling, dtype=np.float32)
        self._y = np.array(y_grid * zernike_scaling, dtype=np.float32)
        self._kernel = None
        self._target_vector = self.spot_amp
        super().__init__(tuple(slm_shape), amp=amp, phase=phase, slm_shape=tuple(slm_shape), **flags)

    def __len__(self):
        return self.spot_amp.size

    # _spots.py:917-948: the target is the N-vector of spot amplitudes
    def _set_target(self, new_target):
        self.target = np.array(self._target_vector, dtype=self.dtype)
        np.abs(self.target, out=self.target)
        self.target *= 1 / gs_oracle.l2norm(self.target)

    def _build_kernel(self):
        """_spots.py:595-636."""
        px, py, c = monomial_table(self.zernike_basis)
        cw = c @ self.spot_zernike                      # (M, N), toolbox/phase.py:905
        x = self._x.astype(np.float64).ravel()
        y = self._y.astype(np.float64).ravel()
        phase = np.zeros((len(self), x.size))
        for m in range(len(px)):
            phase += cw[m][:, None] * (x ** int(px[m]) * y ** int(py[m]))[None, :]
        k = np.exp(1j * phase.astype(np.float32)).astype(self.dtype_complex)
        k /= np.sqrt(k.shape[1])
        return k

    def _forward(self):
        """_build_nearfield (_hologram.py:1000-1011, shape == slm_shape) + _spots.py:767-824 + :951-953."""
        if self._kernel is None:
            self._kernel = self._build_kernel()
        self.nearfield = (self.amp * np.exp(1j * self.phase)).astype(self.dtype_complex)
        nf = np.conj(self.nearfield)
        far = np.matmul(self._kernel, nf.ravel()[:, np.newaxis])[:, 0]
        far = np.conj(far)
        far *= 1 / gs_oracle.l2norm(far)
        self.farfield = far.astype(self.dtype_complex)
        self.amp_ff = np.abs(self.farfield)

    def _inverse(self, extract=True):
        """_spots.py:887-915 + _nearfield_extract (_hologram.py:1026-1036)."""
        nf = np.matmul(self.farfield[np.newaxis, :], self._kernel)[0]
        self.nearfield = nf.reshape(self.shape)
        if extract:
            self.phase = np.arctan2(self.nearfield.imag, self.nearfield.real).astype(self.dtype)

    def _update_weights(self):
        """_spots.py:950-989: "computational" is read as "computational_spot"; feedback = amp_ff."""
        fb = self.flags["feedback"]
        if fb == "computational":
            fb = self.flags["feedback"] = "computational_spot"
        if fb != "computational_spot":
            raise ValueError("Feedback '{}' needs a camera.".format(fb))
        gs_oracle.weight_multiplier_update(self.weights, self.amp_ff, self.target, self.flags["method"],
                                           self.flags, self.dtype)

    def _stat_groups(self, groups):
        """_spots.py:1004-1019: only the experimental group is computed by the reference."""
        return {}
