import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "timeout(seconds): per-test limit (pytest-timeout; a plain marker where the plugin is absent)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle binding (test infrastructure)."""
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation debug harness: the product's per-thread kernel bodies compiled for the CPU."""
    import __graft_entry__ as g
    return ctypes.CDLL(g.build_host_emu())


@pytest.fixture(scope="session")
def make_emu(emu_lib):
    from anerend_b200.capi import DdgiHandle

    def make(params):
        return DdgiHandle(emu_lib, "anerend_emu_ddgi_", params)
    return make


@pytest.fixture(scope="session")
def make_gpu():
    """The product: anerend_cuda_ddgi_* through ctypes.  Fails loudly if the library is missing or no GPU is present."""
    import anerend_b200 as A

    def make(params):
        return A.Ddgi(params)
    return make
