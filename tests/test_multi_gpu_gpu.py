"""Multi-GPU parity on real GPUs (needs >= 2 devices on the box; skipped otherwise): the mix bus
summed over peer memory by the engine's own kernels == an NCCL reduce of the rank-local buses ==
a single-GPU render of the whole bank (<= 1e-6, the summation trees differ)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_bus_equals_nccl_reduce_equals_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "peer_bus_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "PEER_BUS_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
