import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and load the product library; CPU-only hosts can still load it."""
    import __graft_entry__ as g
    g.build()
    import reseek_b200
    return reseek_b200.load_library()


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Port, build_port
    build_port()
    return Port
