import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has a hyphen, so it is imported through importlib)."""
    return importlib.import_module("scann-rust_b200")


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg
