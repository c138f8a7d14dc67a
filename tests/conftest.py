import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc_lib():
    import orc
    return orc.load(False)


@pytest.fixture(scope="session")
def orc_lib32():
    import orc
    return orc.load(True)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; on a box without the built .so this fails loudly (no fallback)."""
    from mars_ode_physics_b200 import capi
    lib = capi.load()
    if os.environ.get("B2P_FAKE_ARENA"):  # scratch/fake_arena_check.py: dry run of the test code itself, no device
        return lib
    if lib.b2p_device_count() <= 0:
        pytest.fail("no CUDA device visible to libmars_ode_b200.so")
    return lib
