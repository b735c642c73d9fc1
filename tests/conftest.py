import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    # (the suite runs against libdrr_test.so -- the product's objects plus the drr_test_* accessors; tests/common.py selects it.
    # The product library, libdrr.so, exports none of them: tests/test_host.py checks both.)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle, libdrr.so and libdrr_test.so once per session (in-tree, git-ignored artefacts)."""
    import __graft_entry__ as g
    g.build(quiet=True)
