"""2-GPU NCCL correctness of the gradient exchange.  Needs two B200s: run with
`gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`; skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision,graph", [("bf16", "1"), ("fp32", "0")])
def test_nccl_grad_exchange_matches_average_of_single_gpu_grads(precision, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, LASR_TEST_PRECISION=precision, LASR_TEST_GRAPH=graph)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "ddp_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_WORKER_OK" in r.stdout
