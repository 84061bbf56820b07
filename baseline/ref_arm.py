"""
Reference arm of bench.py: the UNMODIFIED reference (slmsuite 0.4.1) installed into ``baseline/_ref`` with

    cp -r /root/reference /tmp/refcopy      # /root/reference is read-only and setuptools writes egg-info
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
           --target baseline/_ref /tmp/refcopy

(``baseline/_ref`` is git-ignored and travels to the GPU box with the snapshot).  Its NumPy backend is the
reference's own CPU path (``cp is np``, slmsuite/holography/algorithms/_header.py:16-32).  The package imports
matplotlib and h5py unconditionally (_header.py:1-2, analysis/files.py) -- neither is installed in this image and
neither is touched by ``Hologram.optimize`` -- so empty stub modules are registered before the import; nothing of the
reference is modified.

Only bench.py's ``--impl reference`` / ``cpu_baseline`` legs use this module.
"""
import os
import sys
import types
import warnings

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "slmsuite"))


def load():
    """Returns the reference's ``slmsuite.holography.algorithms`` module (NumPy backend)."""
    if not available():
        raise RuntimeError("baseline/_ref is missing (see baseline/ref_arm.py for the install command)")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "h5py"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ax = sys.modules["mpl_toolkits.axes_grid1"]
    if not hasattr(ax, "make_axes_locatable"):
        ax.make_axes_locatable = lambda *a, **k: None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import slmsuite.holography.algorithms as algorithms
    assert os.path.abspath(algorithms.__file__).startswith(REF_DIR), algorithms.__file__
    return algorithms
