"""GPU suite (-m gpu), multi-GPU part: runs tests/multi_gpu_check.py under torchrun on two GPUs of the box
(skipped on a single-GPU box): in-kernel NVLink exchange of the loss sums == NCCL all-reduce == whole batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_rank_peer_exchange_equals_nccl_and_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(repo, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900, cwd=repo)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    assert "multi_gpu_check ok" in out.stdout
