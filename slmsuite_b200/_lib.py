"""
ctypes binding of ``libslmgs.so`` (C ABI declared in ``include/slmgs.h``).

The product path has exactly one backend: the CUDA library built for sm_100a by
``slmsuite_b200/csrc/Makefile`` (or ``__graft_entry__.build()``).  If it is missing or cannot
be loaded, importing a hologram class works but creating one raises ``RuntimeError`` -- there is
no CPU fallback.  ``use_library(path)`` exists so the CPU test-suite can point the same
bindings at the host *emulation* of the kernel sources (``tests/_emu``, test infrastructure);
nothing in the package calls it.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(_HERE, "libslmgs.so")

_lib = None
_lib_path = None

# slmgs_status, include/slmgs.h
OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_STATE = 0, -1, -2, -3, -4

# slmgs_method: ALGORITHM_DEFAULTS keys of the reference (algorithms/_header.py:53-71)
METHODS = {"GS": 0, "WGS-Leonardo": 1, "WGS-Kim": 2, "WGS-Nogrette": 3, "WGS-Wu": 4, "WGS-tanh": 5}
PHASE_COMPUTE, PHASE_COMPUTE_STORE, PHASE_STORED = 0, 1, 2


class Params(C.Structure):
    """slmgs_params, include/slmgs.h."""

    _fields_ = [
        ("method", C.c_int),
        ("update_weights", C.c_int),
        ("phase_mode", C.c_int),
        ("feedback_exponent", C.c_float),
        ("feedback_factor", C.c_float),
        ("mraf", C.c_int),
        ("mraf_has_factor", C.c_int),
        ("mraf_factor", C.c_float),
        ("feedback", C.c_int),
        ("spot_width", C.c_int),
        ("zero_weights", C.c_int),
        ("zero_factor", C.c_float),
    ]


_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_ctx = C.c_void_p
_pp = C.POINTER(Params)

# name -> (restype, argtypes); every symbol include/slmgs.h declares
SIGNATURES = {
    "slmgs_version": (C.c_int, []),
    "slmgs_last_error": (C.c_char_p, [_ctx]),
    "slmgs_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "slmgs_destroy": (C.c_int, [_ctx]),
    "slmgs_sync": (C.c_int, [_ctx]),
    "slmgs_phase_device_ptr": (C.c_void_p, [_ctx]),
    "slmgs_stream": (C.c_void_p, [_ctx]),
    "slmgs_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "slmgs_set_phase": (C.c_int, [_ctx, _fp]),
    "slmgs_get_phase": (C.c_int, [_ctx, _fp]),
    "slmgs_set_amp_scalar": (C.c_int, [_ctx, C.c_float]),
    "slmgs_set_amp_array": (C.c_int, [_ctx, _fp, C.c_int]),
    "slmgs_set_propagation": (C.c_int, [_ctx, _fp]),
    "slmgs_set_target": (C.c_int, [_ctx, _fp, C.c_int]),
    "slmgs_get_target": (C.c_int, [_ctx, _fp]),
    "slmgs_reset_weights": (C.c_int, [_ctx]),
    "slmgs_set_weights": (C.c_int, [_ctx, _fp]),
    "slmgs_get_weights": (C.c_int, [_ctx, _fp]),
    "slmgs_set_phase_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_get_phase_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_get_amp_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_get_farfield": (C.c_int, [_ctx, _fp]),
    "slmgs_get_phase_gray": (C.c_int, [_ctx, C.c_int, _dp, C.c_void_p]),
    "slmgs_set_sample_grid": (C.c_int, [_ctx, C.c_longlong, _dp, _dp]),
    "slmgs_sample_intensity": (C.c_int, [_ctx, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "slmgs_run": (C.c_int, [_ctx, _pp, C.c_int, C.c_int]),
    "slmgs_set_sparse": (C.c_int, [_ctx, C.c_int]),
    "slmgs_sparse_info": (C.c_int, [_ctx, _ip]),
    "slmgs_forward": (C.c_int, [_ctx]),
    "slmgs_update_weights": (C.c_int, [_ctx, _pp]),
    "slmgs_set_spots": (C.c_int, [_ctx, C.c_int, _ip, _ip, _fp]),
    "slmgs_update_weights_spot": (C.c_int, [_ctx, _pp, C.c_int]),
    "slmgs_constrain_inverse": (C.c_int, [_ctx, _pp]),
    "slmgs_populate": (C.c_int, [_ctx]),
    "slmgs_share_stream": (C.c_int, [_ctx, _ctx]),
    "slmgs_nearfield_sum_ptr": (C.c_void_p, [_ctx]),
    "slmgs_constrain_accumulate": (C.c_int, [_ctx, _pp, C.c_float, C.c_void_p, C.c_int]),
    "slmgs_extract_phase_from_sum": (C.c_int, [_ctx, C.c_void_p]),
    "slmgs_run_accumulate": (C.c_int, [_ctx, _pp, C.c_float, C.c_void_p, C.c_int]),
    "slmgs_stats_pixel": (C.c_int, [_ctx, _dp, _dp]),
    "slmgs_window_power": (C.c_int, [_ctx, C.c_int, _ip, _ip, C.c_int, _dp, _dp]),
    "slmgs_save_phase": (C.c_int, [_ctx]),
    "slmgs_restore_phase": (C.c_int, [_ctx]),
    "slmgs_timer_start": (C.c_int, [_ctx]),
    "slmgs_timer_stop": (C.c_int, [_ctx, _fp]),
    "slmgs_profile_enable": (C.c_int, [_ctx, C.c_int]),
    "slmgs_profile_read": (C.c_int, [_ctx, _fp, _ip]),
    "slmgs_launch_count": (C.c_longlong, [_ctx]),
    "slmgs_launch_geometry": (C.c_int, [_ctx, _ip]),
    "slmgs_time_kernel": (C.c_int, [_ctx, C.c_int, C.c_int, _fp]),
    # compressed spot hologram
    "slmgs_comp_last_error": (C.c_char_p, [_ctx]),
    "slmgs_comp_create": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "slmgs_comp_destroy": (C.c_int, [_ctx]),
    "slmgs_comp_sync": (C.c_int, [_ctx]),
    "slmgs_comp_launch_count": (C.c_longlong, [_ctx]),
    "slmgs_comp_set_basis": (C.c_int, [_ctx, _dp, _dp]),
    "slmgs_comp_set_phase": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_get_phase": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_set_amp_scalar": (C.c_int, [_ctx, C.c_float]),
    "slmgs_comp_set_amp_array": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_set_target": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_set_weights": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_get_weights": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_set_phase_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_get_phase_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_get_amp_ff": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_get_farfield": (C.c_int, [_ctx, _fp]),
    "slmgs_comp_forward": (C.c_int, [_ctx, C.c_int]),
    "slmgs_comp_run": (C.c_int, [_ctx, _pp, C.c_int, C.c_int]),
    "slmgs_comp_near2far": (C.c_int, [_ctx]),
    "slmgs_comp_facc_ptr": (C.c_void_p, [_ctx]),
    "slmgs_comp_stream": (C.c_void_p, [_ctx]),
    "slmgs_comp_constrain_far2near": (C.c_int, [_ctx, _pp]),
    "slmgs_comp_finalize": (C.c_int, [_ctx, C.c_int]),
    "slmgs_comp_timer": (C.c_int, [_ctx, C.c_int, _fp]),
    # multi-GPU collective (NCCL loaded inside the library)
    "slmgs_comm_last_error": (C.c_char_p, []),
    "slmgs_comm_nccl_version": (C.c_int, []),
    "slmgs_comm_unique_id": (C.c_int, [C.c_char_p]),
    "slmgs_comm_create": (C.c_int, [C.POINTER(_ctx), C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "slmgs_comm_destroy": (C.c_int, [_ctx]),
    "slmgs_allgather_phase": (C.c_int, [_ctx, _ctx, C.c_int, C.c_int, C.c_longlong, _fp, C.POINTER(C.c_void_p), _fp]),
    "slmgs_comm_allreduce_f64": (C.c_int, [_ctx, C.c_void_p, C.c_longlong, C.c_void_p]),
}


def _bind(lib):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


def use_library(path):
    """Load a specific build of the C ABI (tests use this for the host emulation)."""
    global _lib, _lib_path
    _lib = _bind(C.CDLL(path))
    _lib_path = path
    return _lib


def library_path():
    return _lib_path


def lib():
    """The loaded C ABI; loads ``libslmgs.so`` next to this file on first use."""
    if _lib is None:
        if not os.path.exists(DEFAULT_LIBRARY):
            raise RuntimeError(
                "slmsuite_b200: CUDA library not built ({} missing). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C slmsuite_b200/csrc`. There is no CPU fallback.".format(DEFAULT_LIBRARY)
            )
        try:
            use_library(DEFAULT_LIBRARY)
        except OSError as exc:
            raise RuntimeError("slmsuite_b200: cannot load {}: {}".format(DEFAULT_LIBRARY, exc)) from exc
    return _lib


def _parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of sysfs cpulist files)."""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device(device):
    """One process per GPU: run this process (and allocate what it pins from now on) on the host cores next to CUDA
    device ``device`` -- the ``local_cpulist`` of the GPU's PCI function in sysfs.  On a two-socket 8-GPU box the
    host <-> device copies of the end-to-end path otherwise cross the socket interconnect for half of the ranks.
    Returns the core list, or None when the topology is not visible (nothing is changed then)."""
    buf = C.create_string_buffer(32)
    try:
        if lib().slmgs_device_pci_bus_id(int(device), buf, 32) != 0 or not buf.value:
            return None
        with open("/sys/bus/pci/devices/{}/local_cpulist".format(buf.value.decode().lower())) as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError, AttributeError):
        return None


class SlmgsError(RuntimeError):
    pass


def check(ctx, status):
    """Map slmgs_status to the exception types the reference raises."""
    if status == OK:
        return
    msg = lib().slmgs_last_error(ctx)
    msg = msg.decode() if msg else "slmgs error {}".format(status)
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_OOM:
        raise MemoryError(msg)
    raise SlmgsError(msg)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def fptr(a):
    return a.ctypes.data_as(_fp)


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)
