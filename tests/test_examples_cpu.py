"""The examples parse and reach the library: without a CUDA device each one must stop at its first compute call with
libpicgolf's "no CPU path" error (and nothing else -- a wrong keyword or import would fail earlier, differently)."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = sorted(glob.glob(os.path.join(ROOT, "examples", "*.py")))


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present: the examples would run in full")
@pytest.mark.parametrize("path", EXAMPLES, ids=[os.path.basename(p) for p in EXAMPLES])
def test_example_stops_at_the_first_compute_call(path, tmp_path):
    r = subprocess.run([sys.executable, path], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode != 0
    assert "no CPU path" in r.stderr and "PicGolfError" in r.stderr, r.stderr[-2000:]


def test_examples_exist():
    assert len(EXAMPLES) == 4
