"""GPU: the N > 1 path on hardware -- shard parity per rank and the per-step statistics exchange in both forms.

One process per rank under torchrun (gloo rendezvous on 127.0.0.1).  With a single GPU the two ranks share it: cudaIpc,
the peer-memory pushes and NCCL all work between two processes on one device, so the multi-rank code runs on the
1-GPU box of the driver as well; with more GPUs visible every rank takes its own."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_two_ranks_shard_parity_and_stats_exchange(mode):
    import torch
    if mode == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("NCCL refuses two ranks on one device")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp_rank_check.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count(" ok (") == 2, r.stdout
