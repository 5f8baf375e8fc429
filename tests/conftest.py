import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "knp-emi-dg_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """host-emulation build of the library sources (CPU test-suite only)"""
    from common import lib_for
    return lib_for("emu")


@pytest.fixture(scope="session")
def gpu_lib():
    """the product library; fails loudly when it is not built or no GPU is there"""
    from common import lib_for
    return lib_for("gpu")
