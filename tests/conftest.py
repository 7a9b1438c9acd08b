import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the oracle is test infrastructure: build it on demand (gcc only, ~1 s)
    so = os.path.join(ROOT, "oracle", "libsgc_oracle.so")
    src = os.path.join(ROOT, "oracle", "sgc_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def golden_path(*p):
    return os.path.join(GOLDEN, *p)


@pytest.fixture(scope="session", autouse=True)
def _library_built():
    """Every test may touch libsgc_b200.so (the host feeder runs without a GPU): build it once per
    session (a no-op when it is up to date; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
