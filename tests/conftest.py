import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def apde_lib():
    """the product library; built in-tree if missing (nvcc cross-compiles without a GPU)"""
    from apde_mvs_b200 import build as _b
    _b.build()
    from apde_mvs_b200.binding import load_library
    return load_library()


@pytest.fixture(scope="session")
def ctx(apde_lib):
    from apde_mvs_b200.binding import Context
    c = Context(0)
    yield c
    c.close()
