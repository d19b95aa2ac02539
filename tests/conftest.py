import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "spectroplot-js_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One engine (== one reference worker) on cuda:0.  No fallback: fails if the CUDA library
    or the device is missing."""
    import spectro_b200
    e = spectro_b200.Engine(0)
    yield e
    e.close()
