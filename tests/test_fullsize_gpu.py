"""Full-size checks on the B200 (BASELINE.json sizes): side-by-side parity with the oracle for a few
iterations, and size-independent properties of the loop where the oracle would take minutes."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _spots(shape, n, seed):
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, shape[0], (2, n))
    t = np.zeros(shape, dtype=np.float32)
    t[pts[1], pts[0]] = 1
    return t


def test_config2_three_iterations_vs_oracle(cuda):
    """BASELINE configs[1] (1152x1920 in 4096^2, WGS-Kim) for 3 iterations against the oracle; 1e-5 rel-RMSE."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    shape, slm = (4096, 4096), (1152, 1920)
    target = _spots(shape, 64, 1)
    phase = np.random.default_rng(2).uniform(-np.pi, np.pi, slm).astype(np.float32)
    kw = dict(method="WGS-Kim", maxiter=3, verbose=False, fix_phase_iteration=2)
    a = Hologram(target, phase=phase, slm_shape=slm)
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleHologram(target, phase=phase, slm_shape=slm)
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert a.flags["fixed_phase"] == b.flags["fixed_phase"] is True
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5


def test_dense_gs_4096_properties(cuda):
    """GS on a dense 4096^2 target: Parseval (ortho transform of a unit-norm near field), efficiency
    non-decreasing over iterations (GS error-reduction property), phase range."""
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(3)
    target = rng.random((4096, 4096), dtype=np.float32)
    h = Hologram(target, phase=rng.uniform(-np.pi, np.pi, (4096, 4096)).astype(np.float32))
    eff = []

    def overlap(holo):
        a = holo.amp_ff
        eff.append(float(np.sum(a.astype(np.float64) * holo.target) ** 2 / np.sum(a.astype(np.float64) ** 2)))

    h.optimize("GS", maxiter=1, verbose=False)
    overlap(h)
    for _ in range(4):
        h.optimize("GS", maxiter=3, verbose=False)
        overlap(h)
    a = h.amp_ff.astype(np.float64)
    assert abs(np.sum(a * a) - 1) < 1e-5
    assert all(e1 >= e0 - 1e-6 for e0, e1 in zip(eff, eff[1:]))
    assert eff[-1] > eff[0]
    ph = h.phase
    assert ph.min() >= -np.pi - 1e-6 and ph.max() <= np.pi + 1e-6


def test_single_delta_gives_blaze_4096(cuda):
    """reference tests/holography/test_algorithms.py:51-84 at 4096^2: one far-field pixel -> linear phase ramp."""
    from slmsuite_b200 import Hologram

    N = 4096
    ky, kx = 1234, 3210
    t = np.zeros((N, N), dtype=np.float32)
    t[ky, kx] = 1
    h = Hologram(t, phase=np.random.default_rng(4).uniform(-np.pi, np.pi, (N, N)).astype(np.float32))
    h.optimize("WGS-Kim", maxiter=8, verbose=False, fix_phase_iteration=3)
    y = np.arange(N, dtype=np.float64)[:, None] - N // 2
    x = np.arange(N, dtype=np.float64)[None, :] - N // 2
    blaze = 2 * np.pi * ((kx - N // 2) * x / N + (ky - N // 2) * y / N)
    err = np.angle(np.exp(1j * (h.get_phase() - blaze)))
    err = np.angle(np.exp(1j * (err - err.flat[0])))
    assert np.allclose(err, 0, atol=1e-2)
    a = h.amp_ff
    assert a[ky, kx] > 0.999


def test_fused_equals_stepped_2048(cuda):
    from slmsuite_b200 import Hologram

    shape, slm = (2048, 2048), (1080, 1920)
    target = _spots(shape, 100, 5)
    phase = np.random.default_rng(6).uniform(-np.pi, np.pi, slm).astype(np.float32)
    a = Hologram(target, phase=phase, slm_shape=slm)
    a.optimize("WGS-Leonardo", maxiter=10, verbose=False)
    b = Hologram(target, phase=phase, slm_shape=slm)
    b.optimize("WGS-Leonardo", maxiter=10, verbose=False, callback=lambda h: False)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5


def test_spot_hologram_config3_runs_and_improves_uniformity(cuda):
    """BASELINE configs[2]: 32x32 spot grid on 4096^2, WGS with per-spot feedback (shortened to 12 iterations)."""
    from slmsuite_b200 import SpotHologram

    h = SpotHologram.make_rectangular_array((4096, 4096), array_shape=(32, 32), array_pitch=(64, 64), basis="knm")
    h.reset_phase(np.random.default_rng(7).uniform(-np.pi, np.pi, (4096, 4096)).astype(np.float32))
    h.optimize("WGS-Leonardo", maxiter=12, verbose=False, feedback="computational_spot",
               stat_groups=["computational_spot"])
    st = h.stats["stats"]["computational_spot"]
    assert st["uniformity"][-1] > st["uniformity"][1]
    assert st["efficiency"][-1] > 0.5


def test_batch_config4_slice(cuda):
    """BASELINE configs[3] shape (2048^2 holograms, GS) with a batch of 4: equals four single holograms."""
    from slmsuite_b200 import Hologram, HologramBatch

    B, N = 4, 2048
    T = np.stack([_spots((N, N), 100, 100 + b) for b in range(B)])
    P = np.stack([np.random.default_rng(200 + b).uniform(-np.pi, np.pi, (N, N)).astype(np.float32) for b in range(B)])
    hb = HologramBatch(T, phase=P)
    hb.optimize("GS", maxiter=6, verbose=False)
    ph = hb.phase
    for b in (0, 3):
        h = Hologram(T[b], phase=P[b])
        h.optimize("GS", maxiter=6, verbose=False)
        assert np.allclose(ph[b], h.phase, atol=1e-5)


def test_8192_config5_shape_runs(cuda):
    """BASELINE configs[4] shape: 8192^2 padded field, 10k random spots, WGS-Leonardo (3 iterations)."""
    from slmsuite_b200 import SpotHologram

    v = np.random.default_rng(5).uniform(64, 8192 - 64, (2, 10000))
    h = SpotHologram((8192, 8192), v, basis="knm")
    h.reset_phase(np.random.default_rng(8).uniform(-np.pi, np.pi, (8192, 8192)).astype(np.float32))
    h.optimize("WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
    a = h.amp_ff.astype(np.float64)
    assert abs(np.sum(a * a) - 1) < 1e-5
    assert h.iter == 3


def test_config2_sparse_equals_dense_at_full_size(cuda):
    """BASELINE configs[1] with the 64-spot target: the sparse far-field path (64 of 2048 column tiles) and the
    dense loop give the same result, 12 WGS-Kim iterations across the phase-fixing iteration."""
    from slmsuite_b200 import Hologram

    shape, slm = (4096, 4096), (1152, 1920)
    target = _spots(shape, 64, 1)
    phase = np.random.default_rng(4).uniform(-np.pi, np.pi, slm).astype(np.float32)
    out = []
    for sparse in (True, False):
        h = Hologram(target, phase=phase, slm_shape=slm)
        h.set_sparse(sparse)
        h.optimize("WGS-Kim", maxiter=12, verbose=False, fix_phase_iteration=5)
        out.append(h)
    a, b = out
    used, n_active, n_tiles = a.sparse_info()
    assert used and n_active <= 64 and n_tiles == 2048
    assert not b.sparse_info()[0]
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 2e-6
    assert rel_rmse(a.weights, b.weights) <= 2e-6
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 2e-5


def test_config4_batch_sparse_per_hologram(cuda):
    """BASELINE configs[3] shard (8 holograms of 2048^2, 100 spots each, GS): per-hologram tile lists vs the dense loop."""
    from slmsuite_b200 import HologramBatch

    B, shape = 8, (2048, 2048)
    targets = np.stack([_spots(shape, 100, 100 + b) for b in range(B)])
    phases = np.random.default_rng(5).uniform(-np.pi, np.pi, (B,) + shape).astype(np.float32)
    res = []
    for sparse in (True, False):
        hb = HologramBatch(targets, phase=phases)
        hb.set_sparse(sparse)
        hb.optimize("GS", maxiter=5, verbose=False)
        res.append(hb)
    a, b = res
    used, n_active, n_tiles = a.sparse_info()
    assert used and n_active <= 100 and n_active < n_tiles // 2
    assert rel_rmse(a.phase, b.phase) <= 1e-6
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-6


def test_compressed_500_spots_vs_oracle(cuda):
    """CompressedSpotHologram at a size the oracle still holds in memory (500 spots on a 256x384 SLM, 3-D, WGS-Kim)."""
    from oracle import compressed_oracle
    from slmsuite_b200 import CompressedSpotHologram

    rng = np.random.default_rng(6)
    slm = (256, 384)
    yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
    grid = ((xx - slm[1] / 2) * 12.6, (yy - slm[0] / 2) * 12.6)
    v = rng.uniform(-0.03, 0.03, (3, 500))
    v[2] = rng.uniform(-2e-5, 2e-5, 500)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    kw = dict(method="WGS-Kim", maxiter=6, verbose=False, fix_phase_iteration=3)
    args = dict(basis="kxy", slm_grid=grid, zernike_scaling=1.0 / 3000.0, phase=phase)
    a = CompressedSpotHologram(v, **args)
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = compressed_oracle.OracleCompressedSpotHologram(v, **args)
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    assert a.flags["fixed_phase"] == b.flags["fixed_phase"] is True
    dphi = np.angle(np.exp(1j * (a.phase.astype(np.float64) - b.phase.astype(np.float64))))
    assert np.sqrt(np.mean(dphi ** 2)) <= 1e-4


def test_camera_2048_vs_oracle(cuda):
    """SimulatedCamera on a 1024^2 SLM in a 2048^2 far field, 1200x1600 camera with a rotated affine grid."""
    from oracle import camera_oracle
    from slmsuite_b200 import SimulatedCamera

    rng = np.random.default_rng(7)
    slm = (1024, 1024)
    yy, xx = np.mgrid[0:1200, 0:1600].astype(np.float64)
    c, s = np.cos(0.1), np.sin(0.1)
    knm = np.array([1024 + 1.6 * (c * (yy - 600) - s * (xx - 800)), 1024 + 1.6 * (s * (yy - 600) + c * (xx - 800))])
    display = rng.integers(0, 256, slm).astype(np.uint8)
    cam = SimulatedCamera(slm, resolution=(1600, 1200), knm_cam=knm, shape_padded=(2048, 2048), bitdepth=12)
    cam.set_exposure(2.0e5)
    img = cam.get_image(display, 256)
    phase = camera_oracle.phase_from_display(display, 256, np.zeros(slm))
    gold, raw = camera_oracle.camera_image(phase, np.ones(slm), slm, (2048, 2048), (1200, 1600), knm, 2.0e5, 1, 12)
    assert img.dtype == gold.dtype == np.uint16
    mine = cam.get_farfield_intensity() * np.float32(2.0e5)
    assert rel_rmse(mine, raw) <= 1e-5
    assert np.array_equal(mine == 0, raw == 0) and (raw == 0).any() and (raw > 0).any()
    diff = img.astype(np.int64) - gold.astype(np.int64)
    near = np.abs(raw - np.rint(raw)) <= 1e-3 * np.maximum(raw, 1.0)
    assert np.all(np.abs(diff) <= 1) and not np.any((diff != 0) & ~near)


# ------------------------------------------------------------------------------------------------------------------
# Full-size parity against the oracle on every BASELINE config.  Tolerance: north_star's 1e-5 rel-RMSE
# of the far-field amplitude; phases 2e-5 rad rms (5e-5 after 30 iterations).  The oracle needs ~5 s per iteration at
# 4096^2 and ~20 s at 8192^2 on one host core, so the iteration counts are small where the size is large.
# ------------------------------------------------------------------------------------------------------------------
def _phase_rms(a, b, mask=None):
    d = np.angle(np.exp(1j * (np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
    if mask is not None:
        d = d[mask]
    return float(np.sqrt(np.mean(d ** 2)))


@pytest.mark.parametrize("variant", ["dense", "delta"])
def test_config1_512_gs_30_iterations_vs_oracle(cuda, variant):
    """BASELINE configs[0] exactly as SURVEY.md 8d writes it: 512^2, default_rng(0), GS, 30 iterations; dense target and
    the reference test's single-delta form (tests/holography/test_algorithms.py:51-84)."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(0)
    if variant == "dense":
        target = rng.random((512, 512), dtype=np.float32)
    else:
        target = np.zeros((512, 512), dtype=np.float32)
        target[rng.integers(512), rng.integers(512)] = 1
    phase = rng.uniform(-np.pi, np.pi, (512, 512)).astype(np.float32)
    a = Hologram(target, phase=phase)
    a.optimize("GS", maxiter=30, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleHologram(target, phase=phase)
        b.optimize("GS", maxiter=30, verbose=False)
    assert a.iter == b.iter == 30
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    nf = np.abs(b.nearfield)
    assert _phase_rms(a.phase, b.phase, nf > 1e-4 * nf.max()) <= 5e-5


def test_dense_gs_4096_three_iterations_vs_oracle(cuda):
    """The metric's configuration (shape == slm_shape == 4096^2, dense target, GS): 3 iterations against the oracle."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(11)
    target = rng.random((4096, 4096), dtype=np.float32)
    phase = rng.uniform(-np.pi, np.pi, (4096, 4096)).astype(np.float32)
    a = Hologram(target, phase=phase)
    a.optimize("GS", maxiter=3, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleHologram(target, phase=phase)
        b.optimize("GS", maxiter=3, verbose=False)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    nf = np.abs(b.nearfield)
    assert _phase_rms(a.phase, b.phase, nf > 1e-4 * nf.max()) <= 2e-5


def test_config3_spot_hologram_4096_vs_oracle(cuda):
    """BASELINE configs[2]: SpotHologram 32x32 on 4096^2, WGS-Leonardo with computational_spot feedback, 3 iterations
    against OracleSpotHologram (_spots.py:1573-1624, analysis/__init__.py:61-204)."""
    from oracle import gs_oracle
    from slmsuite_b200 import SpotHologram

    shape = (4096, 4096)
    phase = np.random.default_rng(12).uniform(-np.pi, np.pi, shape).astype(np.float32)
    kw = dict(method="WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
    a = SpotHologram.make_rectangular_array(shape, array_shape=(32, 32), array_pitch=(64, 64), basis="knm")
    a.reset_phase(phase)
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleSpotHologram.make_rectangular_array(shape, array_shape=(32, 32), array_pitch=(64, 64), basis="knm")
        b.reset_phase(phase)
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5
    nf = np.abs(b.nearfield)
    assert _phase_rms(a.phase, b.phase, nf > 1e-4 * nf.max()) <= 2e-5


@pytest.mark.slow
def test_config5_8192_10k_spots_vs_oracle(cuda):
    """BASELINE configs[4]: SpotHologram 10k random spots on 8192^2, WGS-Leonardo with computational_spot feedback,
    2 iterations against the oracle (about a minute of host time)."""
    from oracle import gs_oracle
    from slmsuite_b200 import SpotHologram

    shape = (8192, 8192)
    v = np.random.default_rng(5).uniform(64, 8192 - 64, (2, 10000))
    phase = np.random.default_rng(13).uniform(-np.pi, np.pi, shape).astype(np.float32)
    kw = dict(method="WGS-Leonardo", maxiter=2, verbose=False, feedback="computational_spot")
    a = SpotHologram(shape, v, basis="knm")
    a.reset_phase(phase)
    a.optimize(**kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = gs_oracle.OracleSpotHologram(shape, v, basis="knm")
        b.reset_phase(phase)
        b.optimize(**kw)
    assert rel_rmse(a.amp_ff, b.amp_ff) <= 1e-5
    assert rel_rmse(a.weights, b.weights) <= 1e-5


def test_config4_batch_member_vs_oracle(cuda):
    """BASELINE configs[3]: one hologram of a 2048^2 GS batch (100 unit spots, default_rng(100 + b) / (200 + b)) against
    the ORACLE, not against this library's own single-hologram path."""
    from oracle import gs_oracle
    from slmsuite_b200 import HologramBatch

    B, N = 4, 2048
    T = np.stack([_spots((N, N), 100, 100 + b) for b in range(B)])
    P = np.stack([np.random.default_rng(200 + b).uniform(-np.pi, np.pi, (N, N)).astype(np.float32) for b in range(B)])
    hb = HologramBatch(T, phase=P)
    hb.optimize("GS", maxiter=5, verbose=False)
    amp, ph = hb.amp_ff, hb.phase
    for b in (1, 3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = gs_oracle.OracleHologram(T[b], phase=P[b])
            o.optimize("GS", maxiter=5, verbose=False)
        assert rel_rmse(amp[b], o.amp_ff) <= 1e-5
        nf = np.abs(o.nearfield)
        assert _phase_rms(ph[b], o.phase, nf > 1e-4 * nf.max()) <= 2e-5


@pytest.mark.parametrize("case", ["dense_gs", "dense_kim", "padded_kim", "padded_gs"])
def test_team_kernels_bit_identical_to_plain_kernels(cuda, case, monkeypatch):
    """The TMA / two-team kernels of the dense-far-field loop (csrc/slmgs_teams.h) call the same arithmetic as the
    plain fused kernels: SLMGS_TEAMS=1 and SLMGS_TEAMS=0 must agree BIT FOR BIT (phase, weights, far-field
    amplitude) on the 4096^2 configurations, dense and zero padded, GS and WGS-Kim across the phase-fixing iteration."""
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(11)
    shape = (4096, 4096)
    slm = shape if case.startswith("dense") else (1152, 1920)
    target = rng.random(shape, dtype=np.float32)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    method = "GS" if "gs" in case else "WGS-Kim"
    kw = dict(method=method, maxiter=6, verbose=False)
    if method == "WGS-Kim":
        kw["fix_phase_iteration"] = 3
    out = []
    for teams in ("1", "0"):
        monkeypatch.setenv("SLMGS_TEAMS", teams)
        monkeypatch.setenv("SLMGS_SPARSE", "0")
        h = Hologram(target, phase=phase, slm_shape=slm)
        h.optimize(**kw)
        out.append((h.phase.copy(), h.weights.copy(), h.amp_ff.copy()))
        del h
    for a, b in zip(*out):
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("shape,batch", [((2048, 2048), 3), ((4096, 4096), 2), ((2048, 4096), 2), ((4096, 2048), 2)])
def test_team_kernels_batch_bit_identical(cuda, shape, batch, monkeypatch):
    """Batched contexts (blockIdx.y = hologram, the persistent blocks of a team kernel are shared by the batch) and
    rectangular 2048 / 4096 shapes (team row kernel at 2048-point rows, team column kernel only at 4096-point columns):
    SLMGS_TEAMS=1 and SLMGS_TEAMS=0 agree bit for bit."""
    from slmsuite_b200 import HologramBatch

    rng = np.random.default_rng(21)
    targets = rng.random((batch,) + shape, dtype=np.float32)
    phases = rng.uniform(-np.pi, np.pi, (batch,) + shape).astype(np.float32)
    out = []
    for teams in ("1", "0"):
        monkeypatch.setenv("SLMGS_TEAMS", teams)
        monkeypatch.setenv("SLMGS_SPARSE", "0")
        hb = HologramBatch(targets, phase=phases)
        hb.optimize("WGS-Leonardo", maxiter=4, verbose=False)
        out.append((hb.phase.copy(), hb.amp_ff.copy()))
        del hb
    for a, b in zip(*out):
        assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n,slm", [(512, (512, 512)), (512, (300, 200)), (1024, (1024, 1024)), (256, (256, 256))])
def test_loop_kernel_bit_identical(cuda, n, slm, monkeypatch):
    """Small square fields run the whole GS loop in one cooperative kernel (csrc/slmgs_loop.h: the phases of the plain
    kernels between grid barriers).  SLMGS_LOOP=1 and SLMGS_LOOP=0 agree bit for bit, and the loop kernel is really
    used (far fewer launches)."""
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(31)
    target = rng.random((n, n), dtype=np.float32)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    out = []
    for loop in ("1", "0"):
        monkeypatch.setenv("SLMGS_LOOP", loop)
        monkeypatch.setenv("SLMGS_SPARSE", "0")
        h = Hologram(target, phase=phase, slm_shape=slm)
        h.optimize("GS", maxiter=12, verbose=False)
        out.append((h.phase.copy(), h.amp_ff.copy(), h._lib.slmgs_launch_count(h._ctx)))
        del h
    assert out[0][0].tobytes() == out[1][0].tobytes()
    assert out[0][1].tobytes() == out[1][1].tobytes()
    assert out[0][2] < out[1][2] - 10


@pytest.mark.parametrize("slm", [(4096, 4096), (3000, 3500)])
def test_team_kernels_amp_array_and_propagation_kernel(cuda, slm, monkeypatch):
    """Team row kernel with a per-pixel source amplitude and a propagation kernel (the general projection path), dense and
    zero padded with an odd-sized crop: bit-identical to the plain kernels, and within 1e-5 of the oracle."""
    from oracle import gs_oracle
    from slmsuite_b200 import Hologram

    rng = np.random.default_rng(41)
    shape = (4096, 4096)
    target = rng.random(shape, dtype=np.float32) + 0.1
    yy, xx = np.mgrid[0:slm[0], 0:slm[1]]
    amp = np.exp(-(((xx - slm[1] / 2) / (0.4 * slm[1])) ** 2 + ((yy - slm[0] / 2) / (0.4 * slm[0])) ** 2)).astype(np.float32)
    prop = (1e-6 * ((xx - slm[1] / 2) ** 2 + (yy - slm[0] / 2) ** 2)).astype(np.float32)
    phase = rng.uniform(-np.pi, np.pi, slm).astype(np.float32)
    kw = dict(method="GS", maxiter=3, verbose=False)
    out = []
    for teams in ("1", "0"):
        monkeypatch.setenv("SLMGS_TEAMS", teams)
        monkeypatch.setenv("SLMGS_SPARSE", "0")
        h = Hologram(target, amp=amp, phase=phase, slm_shape=slm, propagation_kernel=prop)
        h.optimize(**kw)
        out.append((h.phase.copy(), h.amp_ff.copy()))
        del h
    for a, b in zip(*out):
        assert a.tobytes() == b.tobytes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = gs_oracle.OracleHologram(target, amp=amp, phase=phase, slm_shape=slm, propagation_kernel=prop)
        ref.optimize(**kw)
    assert rel_rmse(out[0][1], ref.amp_ff) <= 1e-5


@pytest.mark.parametrize("case", ["gs_dense_target", "kim_spots", "spot_feedback"])
def test_8192_split_column_team_kernels_vs_plain(cuda, case, monkeypatch):
    """The opt-in 8192-point team column kernels (csrc/slmgs_teams.h ColKernelT8: a column as two interleaved
    4096-point lines joined by one radix-2 step between lane pairs, SLMGS_TEAMS8=1) use a different factorisation than
    the plain 8192 kernels, so they agree to rounding, not bit for bit: fused GS, the in-kernel power-law update across
    the phase-fixing iteration, and the forward pre-pass of the per-spot feedback."""
    from slmsuite_b200 import Hologram, SpotHologram

    n = 8192
    rng = np.random.default_rng(21)
    phase = rng.uniform(-np.pi, np.pi, (n, n)).astype(np.float32)
    out = []
    for teams8 in ("1", "0"):
        monkeypatch.setenv("SLMGS_TEAMS8", teams8)
        monkeypatch.setenv("SLMGS_SPARSE", "0")
        if case == "gs_dense_target":
            h = Hologram(np.random.default_rng(22).random((n, n), dtype=np.float32), phase=phase, slm_shape=(n, n))
            h.optimize("GS", maxiter=2, verbose=False)
        elif case == "kim_spots":
            h = Hologram(_spots((n, n), 2000, 23), phase=phase, slm_shape=(n, n))
            h.optimize("WGS-Kim", maxiter=5, verbose=False, fix_phase_iteration=3)
        else:
            v = np.random.default_rng(5).uniform(64, n - 64, (2, 2000))
            h = SpotHologram((n, n), v, basis="knm")
            h.reset_phase(phase)
            h.optimize("WGS-Leonardo", maxiter=3, verbose=False, feedback="computational_spot")
        out.append((h.amp_ff.copy(), np.asarray(h.weights).copy(), h.phase.copy()))
        del h
    (amp1, w1, ph1), (amp0, w0, ph0) = out
    assert rel_rmse(amp1, amp0) <= 1e-5
    assert rel_rmse(w1, w0) <= 1e-5
    d = np.angle(np.exp(1j * (ph1.astype(np.float64) - ph0)))
    assert np.sqrt(np.mean(d * d)) <= 2e-5
