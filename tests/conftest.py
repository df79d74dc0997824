import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def nvtt():
    import nvtt_b200_loader
    return nvtt_b200_loader.load()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference built with the pinned parity flags (oracle/_ref, test infrastructure)."""
    import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libnvtt_ref.so not built (needs /root/reference at build time)")
    return refapi


@pytest.fixture(scope="session")
def ctx(nvtt):
    c = nvtt.Context(0)  # raises loudly when the CUDA library or a GPU is missing
    yield c
    c.close()
