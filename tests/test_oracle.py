"""
Pins ``oracle/gs_oracle.py`` (the NumPy restatement) against the golden vectors that
``oracle/make_golden.py`` recorded from the UNMODIFIED reference, and - when the
reference tree is present (build container only) - against the live reference.
CPU only.
"""
import json
import os
import warnings

import numpy as np
import pytest

from oracle import cases, gs_oracle, ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def test_manifest_lists_every_case():
    man = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    assert set(man["cases"]) == set(cases.CASES) | set(cases.MULTI_CASES)
    # the generator recorded a bit-exact restatement on its NumPy
    assert all(v["oracle_max_abs_diff"] == 0.0 for v in man["cases"].values())


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_golden(name):
    gold = _load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = cases.summarize(cases.run_case(name, gs_oracle.OracleHologram, gs_oracle.OracleSpotHologram))
    assert set(got) == set(gold)
    for k in gold:
        # Same NumPy => bit-exact; a different NumPy/pocketfft build may differ in the last ulps,
        # which GS amplifies slowly: allow 2e-4 abs on phases/amplitudes there.
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(gold[k], dtype=np.float64),
                                   rtol=2e-4, atol=2e-4, equal_nan=True, err_msg=f"{name}:{k}")


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", ["spots20_64_WGS-Kim", "padded_kim_128", "spot_rect_padded_128_kim_spotfb",
                                  "mraf_factor_leonardo_64"])
def test_oracle_bit_exact_vs_live_reference(name):
    ref = ref_loader.load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = cases.summarize(cases.run_case(name, ref.Hologram, ref.SpotHologram))
        b = cases.summarize(cases.run_case(name, gs_oracle.OracleHologram, gs_oracle.OracleSpotHologram))
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=f"{name}:{k}")


@pytest.mark.parametrize("name", sorted(cases.MULTI_CASES))
def test_multiplane_oracle_matches_golden(name):
    gold = _load(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = cases.summarize_multi(cases.run_multi_case(name, gs_oracle.OracleHologram, gs_oracle.OracleSpotHologram,
                                                         gs_oracle.OracleMultiplaneHologram))
    assert set(got) == set(gold)
    for k in gold:
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(gold[k], dtype=np.float64),
                                   rtol=2e-4, atol=2e-4, equal_nan=True, err_msg=f"{name}:{k}")


def test_crop_bounds_matches_reference_pad_tests():
    # reference tests/holography/test_toolbox.py:757-830
    assert gs_oracle.crop_bounds((7, 10), (3, 4)) == (2, 5, 3, 7)
    assert gs_oracle.crop_bounds((3, 4), (2, 3)) == (0, 2, 0, 3)
    assert gs_oracle.crop_bounds((4096, 4096), (1152, 1920)) == (1472, 2624, 1088, 3008)
    with pytest.raises(ValueError, match="too small"):
        gs_oracle.crop_bounds((3, 4), (10, 10))


def test_take_sum_matches_reference_take_tests():
    # reference tests/holography/test_analysis.py:591-737: ones over a 10x10 window integrate to 100
    img = np.ones((64, 64), dtype=np.float32)
    out = gs_oracle.take_sum(img, np.array([[20, 30], [20, 40]]), 10)
    assert out.shape == (2,) and np.all(out == 100)
    assert out.dtype == np.float64


def test_padded_shape():
    # _hologram.py:713-723
    assert gs_oracle.padded_shape((1152, 1920), 1) == (2048, 2048)
    assert gs_oracle.padded_shape((1152, 1920), 2) == (4096, 4096)
    assert gs_oracle.padded_shape((720, 1280), 1, square_padding=False) == (1024, 2048)
