"""N-GPU == 1-GPU parity (NCCL all-reduce of the integer charge grid).  Needs >= 2 visible GPUs; skipped otherwise.
The check itself lives in tools/multigpu_check.py and runs under torchrun, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one(pg):
    n = pg.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MULTIGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
