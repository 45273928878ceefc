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


def pytest_collection_modifyitems(config, items):
    """The long full-resolution parity tests run last: the driver runs the suite with -x, and an early stop in one
    of them must not hide the short tests of the other rows."""
    items.sort(key=lambda it: 1 if ("full_size" in it.name or "full_case" in getattr(it, "fixturenames", ())) else 0)
