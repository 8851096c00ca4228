import os
import sys

import pytest

# the single-GPU virtual-rank tests run several ranks' streams (3 per rank) on one device: their device-side waits need
# the other ranks' kernels to run, so no two of those streams may share a hardware queue (default: 8 queues)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and no kernel may be loaded lazily while another rank's wait kernel spins (lazy module loading synchronises the
# context: the host would block inside a launch of rank 0 before it has enqueued rank 1)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The in-tree CUDA library and the CPU emulator must exist before any test runs."""
    import __graft_entry__ as g
    g.build()
    yield
