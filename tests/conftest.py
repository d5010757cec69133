import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def emu_lib():
    """The product sources compiled against the SIMT emulator: same C ABI, executed on the CPU.
    Test infrastructure only -- see tests/emu/simt.h."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from rtk_b200 import api
    lib = api.Library(build_emu.build())
    assert lib.rtk_cuda_init(0) == 0
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real device.  Fails loudly (no skip, no fallback) if the CUDA
    extension is missing or no device is usable."""
    from rtk_b200 import api
    lib = api.load()
    r = lib.rtk_cuda_init(0)
    assert r == 0, "rtk_cuda_init failed: " + lib.last_error()
    return lib
