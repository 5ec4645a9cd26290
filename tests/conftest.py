import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import pyoracle
    pyoracle.build(ref=True)
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx():
    import pcg_b200
    ctx = pcg_b200.Context(0)
    yield ctx
    ctx.close()
