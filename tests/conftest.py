import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "synthetic-to-real-semantic-segmentation_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sub(name=""):
    """Import a sub-module of the (hyphenated) product package."""
    return importlib.import_module(PKG + ("." + name if name else ""))


@pytest.fixture(scope="session")
def pkg():
    return sub()


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as ge
    p = sub("_lib")
    if not os.path.exists(p.LIB_PATH):
        ge.build()
    return p.lib()


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
