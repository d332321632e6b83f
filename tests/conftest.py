import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Build (if stale) the product libraries and the oracle once per session; nvcc cross-compiles on CPU."""
    from lens_trace_b200 import build
    import lt_oracle

    build.build_all()
    lt_oracle.build()
    return True


@pytest.fixture(scope="session")
def root():
    return ROOT
