import os
import subprocess
import sys

import pytest

# several casters of one process meet at device-side barriers in the virtual-rank tests: their streams must not share a
# hardware queue (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and no kernel may be loaded lazily while another caster of the process spins in a barrier kernel (loading a module
# waits for the device); one process per GPU, the way the library is deployed, has no such coupling
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds (if needed) and returns the path of the CPU oracle. Test infrastructure only."""
    out = os.path.join(ROOT, "oracle", "_build", "libmv_oracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".h"))]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return out


@pytest.fixture(scope="session")
def product_lib():
    """Path of the built product library; building it is __graft_entry__.build()'s job."""
    from multivolumes_b200 import LIB_PATH
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "multivolumes_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH
