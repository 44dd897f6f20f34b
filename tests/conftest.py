import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The in-tree build (libmfkc.so, mfkc_cli, the C oracle).  On the GPU box the prebuilt
    files travel with the snapshot; here they are (re)built on demand."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_c(built):
    from tests import _oracle_c
    return _oracle_c.load()


def has_gpu() -> bool:
    try:
        import metafast_b200 as m
        return m.load().mfkc_device_count() > 0
    except Exception:
        return False
