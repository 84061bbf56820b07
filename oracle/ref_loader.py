"""
Import the UNMODIFIED reference (slmsuite) from /root/reference  --  TEST INFRASTRUCTURE ONLY.

The reference tree exists only in the build container, never on the GPU box, so
nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is
used by ``oracle/make_golden.py`` (fixture generation) and by the CPU-only tests
that pin ``oracle/gs_oracle.py`` against the live reference when it is present.

The reference imports matplotlib and h5py unconditionally
(slmsuite/holography/algorithms/_header.py:1-2, analysis/files.py); neither is
installed here and neither is touched by the compute path, so empty stub modules
are registered before the import.
"""

import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "slmsuite"))


def load_reference():
    """Returns the reference's ``slmsuite.holography.algorithms`` module."""
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ax = sys.modules["mpl_toolkits.axes_grid1"]
    if not hasattr(ax, "make_axes_locatable"):
        ax.make_axes_locatable = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import slmsuite.holography.algorithms as algorithms
    return algorithms
