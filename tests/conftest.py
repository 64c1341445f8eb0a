import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ctx():
    import gemma_b200
    c = gemma_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session", autouse=True)
def _fresh_cli():
    """The gemma-b200 binary is a build product (not tracked): (re)build it once per session from the sources so
    that a stale executable can never pass tests the source would fail."""
    import subprocess
    host = os.path.join(ROOT, "gemma_b200", "host")
    if os.path.exists(os.path.join(ROOT, "gemma_b200", "csrc", "libgemma_b200.so")):
        subprocess.check_call(["make", "-s", "-C", host])
    yield
