import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def zkw():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("webauthn-halo2_b200")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure)."""
    from oracle import cpu
    cpu.lib()
    return cpu


@pytest.fixture(scope="session")
def ctx(zkw):
    """One device context for the whole GPU session; fails loudly when the extension or GPU is missing."""
    c = zkw.Context(0)
    yield c
    c.close()
