import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import portlib
    portlib.lib()
    return portlib


@pytest.fixture(scope="session")
def ref():
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/liboracle_usearch.so not built")
    return reflib


@pytest.fixture(scope="session")
def eng():
    from lantern_b200 import api
    api.lib()
    if api.device_count() < 1:
        pytest.fail("no CUDA device: the engine has no CPU fallback")
    return api
