import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "3d-object-detection-of-industrial-joints_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def orc():
    from oracle import pcl_oracle
    pcl_oracle.lib()
    return pcl_oracle


@pytest.fixture(scope="session")
def b200(pkg):
    """The CUDA library binding.  GPU tests must fail (not skip) when the extension is missing."""
    binding = importlib.import_module(PKG_NAME + ".binding")
    binding.lib()
    return binding
