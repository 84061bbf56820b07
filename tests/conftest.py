"""
Test configuration.

Two builds of the same kernel sources can sit behind the C ABI:
  * ``cuda``: slmsuite_b200/libslmgs.so, the product (sm_100a).  Tests that use it are marked
    ``gpu`` and run on the B200 box.
  * ``emu`` : tests/_emu/libslmgs_emu.so, the HOST EMULATION of the very same phase-structured
    kernel sources compiled with g++ -DSLMGS_EMULATE.  Test infrastructure only (it lets the CPU
    suite exercise the index math, the state machine and the host classes); the package never
    loads it on its own.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CSRC = os.path.join(ROOT, "slmsuite_b200", "csrc")
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "libslmgs_emu.so")
CUDA_LIB = os.path.join(ROOT, "slmsuite_b200", "libslmgs.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: long-running")


def _build_emu():
    newest_src = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                     if f.endswith((".h", ".cu")))
    if os.path.exists(EMU_LIB) and os.path.getmtime(EMU_LIB) >= newest_src:
        return
    subprocess.run(["make", "-s", "-j8", "-C", CSRC, "emu"], check=True)


def _build_cuda():
    """libslmgs.so is a build artefact (git-ignored): build it when a fresh checkout has none."""
    if os.path.exists(CUDA_LIB):
        return
    subprocess.run(["make", "-s", "-j8", "-C", CSRC], check=True)


def gpu_present():
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


@pytest.fixture(scope="session")
def emu_library():
    _build_emu()
    return EMU_LIB


@pytest.fixture(scope="session")
def cuda_library():
    _build_cuda()
    return CUDA_LIB


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, emu_library):
    """Loads the requested build behind slmsuite_b200._lib and returns its name."""
    from slmsuite_b200 import _lib

    if request.param == "emu":
        _lib.use_library(emu_library)
    else:
        if not gpu_present():
            pytest.skip("no CUDA device")
        _build_cuda()
        _lib.use_library(CUDA_LIB)
    return request.param


@pytest.fixture
def emu(emu_library):
    from slmsuite_b200 import _lib

    _lib.use_library(emu_library)
    return "emu"


@pytest.fixture
def cuda():
    from slmsuite_b200 import _lib

    if not gpu_present():
        pytest.skip("no CUDA device")
    _build_cuda()
    _lib.use_library(CUDA_LIB)
    return "cuda"
