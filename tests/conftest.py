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
def _gpu_ctx_session():
    import pcg_b200
    ctx = pcg_b200.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture
def gpu_ctx(_gpu_ctx_session):
    """The session's context with the PCGDO_* switches re-read: the library reads the environment once per context,
    tests flip the variables (monkeypatch) and call ``reload_env`` themselves after each change."""
    _gpu_ctx_session.reload_env()
    yield _gpu_ctx_session
    _gpu_ctx_session.reload_env()
