import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the native libraries once per session."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "acados_b200", "csrc", "libcuipm.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_ipm.so")):
        g.build()
    return True
