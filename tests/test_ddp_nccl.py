"""-m gpu, needs >= 2 GPUs (gpurun --gpus 2): NCCL gradient exchange on hardware.  See tests/ddp_nccl_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_mean_gradients_world2():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "ddp_nccl_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-4000:])
    sys.stderr.write(p.stderr[-4000:])
    assert p.returncode == 0, "NCCL worker failed"
    assert "RANK 0: OK" in p.stdout and "RANK 1: OK" in p.stdout
