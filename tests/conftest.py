import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def edb():
    return importlib.import_module("dmft-ed_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ed_oracle
    ed_oracle.build()
    return ed_oracle
