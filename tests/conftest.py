import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture
def fake_kernels(monkeypatch):
    """Host-logic tests: route the kernel launchers to the TEST-ONLY torch-CPU stand-ins."""
    from tests import fake_backend
    import ctgan_b200.tflib as lib
    lib.delete_all_params()
    fake_backend.install(monkeypatch)
    yield
    lib.delete_all_params()
