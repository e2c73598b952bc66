import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; a no-op marker when the plugin is absent)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context of the product library for the whole GPU session."""
    from imagestitch_b200 import build as B, stitching as S
    B.build()
    c = S.Context(0)
    yield c
    c.close()
