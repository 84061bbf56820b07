"""
Generate tests/golden/*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz + MANIFEST.json

For every case in ``oracle/cases.py`` the reference's own ``Hologram`` /
``SpotHologram`` (NumPy backend) is built and optimised, and the final ``phase``,
``amp_ff``, ``weights``, ``iter``, ``fixed_phase`` and per-iteration stats are stored.
The same case is run through ``oracle.gs_oracle`` and the max abs difference is
printed and recorded in the manifest (0.0 everywhere = bit-exact restatement).
"""

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cases, gs_oracle, ref_loader  # noqa: E402


def main():
    ref = ref_loader.load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"numpy": np.__version__, "reference": "slmsuite 0.4.1 @ 39243f08", "cases": {}}
    for name in cases.CASES:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h_ref = cases.run_case(name, ref.Hologram, ref.SpotHologram)
            h_or = cases.run_case(name, gs_oracle.OracleHologram, gs_oracle.OracleSpotHologram)
        g = cases.summarize(h_ref)
        o = cases.summarize(h_or)
        worst = 0.0
        for k in g:
            a, b = np.atleast_1d(np.asarray(g[k], dtype=np.float64)), np.atleast_1d(np.asarray(o[k], dtype=np.float64))
            if a.shape != b.shape:
                worst = float("inf")
                continue
            d = np.abs(a - b)
            d[np.isnan(a) & np.isnan(b)] = 0
            worst = max(worst, float(np.max(d)) if d.size else 0.0)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **g)
        sha = hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
        manifest["cases"][name] = {"oracle_max_abs_diff": worst, "sha256_16": sha,
                                   "keys": sorted(g.keys())}
        print(f"{name:45s} oracle-vs-reference max|d| = {worst:.3e}")
    from slmsuite.holography.algorithms import MultiplaneHologram as RefMulti
    for name in cases.MULTI_CASES:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g = cases.summarize_multi(cases.run_multi_case(name, ref.Hologram, ref.SpotHologram, RefMulti))
            o = cases.summarize_multi(cases.run_multi_case(name, gs_oracle.OracleHologram, gs_oracle.OracleSpotHologram,
                                                           gs_oracle.OracleMultiplaneHologram))
        worst = 0.0
        for k in g:
            a, b = np.atleast_1d(np.asarray(g[k], dtype=np.float64)), np.atleast_1d(np.asarray(o[k], dtype=np.float64))
            worst = max(worst, float(np.max(np.abs(a - b))) if a.shape == b.shape else float("inf"))
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **g)
        sha = hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
        manifest["cases"][name] = {"oracle_max_abs_diff": worst, "sha256_16": sha, "keys": sorted(g.keys())}
        print(f"{name:45s} oracle-vs-reference max|d| = {worst:.3e}")
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
