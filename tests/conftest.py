import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "fmri-connectivity-super-resolution_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_pkg():
    """The package directory name carries hyphens (it is the reference's repo
    name + ``_b200``), so it is imported through importlib."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return _load


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_pkg(pkg):
    """GPU tests go through the product path only; it must fail loudly (not
    skip) when the CUDA library is missing on a GPU box."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    pkg.lib()          # raises if libgsr_sm100.so is missing
    return pkg
