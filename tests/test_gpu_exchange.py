"""GPU parity, two ranks: the exchange of the shard-local hits inside the merge kernel (vs_exchange_*, CUDA IPC + flag
words in peer memory) gives the oracle's top-k, like the NCCL all-gather form.  Needs two GPUs (skipped on one)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_exchange_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "exchange_check.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"mismatches_vs_oracle": 0' in r.stdout
