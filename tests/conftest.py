import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.oracle()   # builds liboracle.so if missing
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_ctx():
    import sjpeg_b200
    ctx = sjpeg_b200.Context(0)    # raises without a device or without the built library
    yield ctx
    ctx.close()
