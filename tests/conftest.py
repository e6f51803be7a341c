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


# ---- parity ledger --------------------------------------------------------------------------------------------
# Every `relnorm(got, want) < bar` a test evaluates is recorded as {test, check, measured, bar}: the measured error next to
# the bar it was held to.  A GPU session writes the ledger to gpurun_out/PARITY.json (the directory that travels back from
# the GPU box); the copy committed at the repo root is what __graft_entry__.smoke() summarises.
_LEDGER = []
_CURRENT = {"test": None}


class Measured(float):
    """A float that remembers the comparison it takes part in."""

    def __new__(cls, value, what):
        obj = super().__new__(cls, value)
        obj.what = what
        return obj

    def _log(self, bar, ok):
        _LEDGER.append({"test": _CURRENT["test"], "check": self.what, "measured": float(self), "bar": float(bar), "ok": bool(ok)})
        return ok

    def __lt__(self, bar):
        return self._log(bar, float(self) < float(bar))

    def __le__(self, bar):
        return self._log(bar, float(self) <= float(bar))


def _caller_line():
    import inspect
    import linecache

    f = inspect.currentframe().f_back.f_back
    line = linecache.getline(f.f_code.co_filename, f.f_lineno).strip()
    return f"{os.path.basename(f.f_code.co_filename)}:{f.f_lineno}: {line[:160]}"


def relnorm(a, b):
    """max|a-b| / max|b|  (norm-wise; element-wise rtol is meaningless at zero crossings)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return Measured(np.abs(a - b).max() / (den if den > 0 else 1.0), _caller_line())


@pytest.fixture(autouse=True)
def _ledger_test_name(request):
    _CURRENT["test"] = request.node.nodeid
    yield
    _CURRENT["test"] = None


def pytest_sessionfinish(session, exitstatus):
    rows = [r for r in _LEDGER if r["test"] and ("gpu" in r["test"] or "parity" in r["test"] or "esfield_gpu" in r["test"])]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not rows or not os.path.isdir(out_dir) or not (os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")):
        return
    import json

    worst = {}
    for r in rows:
        key = (r["test"], r["check"], r["bar"])  # a source line may hold several comparisons with different bars
        if key not in worst or r["measured"] / max(r["bar"], 1e-300) > worst[key]["measured"] / max(worst[key]["bar"], 1e-300):
            worst[key] = r
    with open(os.path.join(out_dir, "PARITY.json"), "w") as fh:
        json.dump({"what": "every relnorm(got, want) < bar evaluated by `pytest -m gpu` on a B200: measured norm-wise error and the bar it was held to "
                           "(worst occurrence per source line); got = CUDA path through the C ABI, want = CPU oracle or committed golden fixture",
                   "checks": sorted(worst.values(), key=lambda r: (r["test"], r["check"], r["bar"]))}, fh, indent=1)
