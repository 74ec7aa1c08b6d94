"""Multi-GPU (NCCL) parity of the env-sharded learner: needs >= 2 GPUs on the box (skipped otherwise; the
CPU-side logic is covered by test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_update_matches_single_process():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
