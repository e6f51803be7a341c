import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests need a device.  On a box WITH a device a missing libpicgolf.so is an error, not a skip.
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                import particleincellcodegolf.jl_b200 as pg
                have = os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")
                if have:
                    have = pg.device_count() > 0
            if not have:
                item.add_marker(pytest.mark.skip(reason="no CUDA device visible"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def pg():
    import particleincellcodegolf.jl_b200 as pg
    return pg


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def relnorm(a, b):
    """max|a-b| / max|b|  (norm-wise; element-wise rtol is meaningless at zero crossings)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)
