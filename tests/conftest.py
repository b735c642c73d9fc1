import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and libdrr.so once per session (both are in-tree, git-ignored artefacts)."""
    import __graft_entry__ as g
    g.build(quiet=True)
