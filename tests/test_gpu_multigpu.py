"""Two-GPU parity under pytest: launches tests/multigpu_check.py with torchrun when the box has >= 2 devices.

multigpu_check.py compares (1) the SUM of the ranks' shard gradients with the single-GPU gradient of the whole batch and
with the CPU oracle, (2) the peer-memory trainer (reduce-scatter + Adam + all-gather in one kernel, plx_adam_step_peer)
with the NCCL all-reduce trainer after 4 steps, and checks that all replicas are bit-equal.  Skipped on one-GPU boxes
(the gloo world_size-2 tests in test_distributed_gloo.py cover the host-side sharding there).
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("fused", ["0", "1"])
def test_two_gpu_gradient_sum_and_peer_trainer(plx_lib, fused):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PLX_PEER_FUSED=fused)      # 1 = cross-GPU waits / signals inside K12 and K3p (PlxPeerSync)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, cwd=REPO, env=env, capture_output=True, text=True, timeout=300)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-3000:]
    assert "MULTIGPU_OK" in text and "PEER_OK" in text, text[-3000:]
