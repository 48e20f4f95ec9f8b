"""Two-GPU parity under pytest: launches tests/multigpu_check.py with torchrun when the box has >= 2 devices.

multigpu_check.py compares (1) the SUM of the ranks' shard gradients with the single-GPU gradient of the whole batch and
with the CPU oracle, (2) the peer-memory trainer of every exchange variant — pull (k_adam_peer), pull with the cross-GPU
waits fused into the kernels, pull through the NVSwitch (k_adam_mc: multimem.ld_reduce / multimem.st, forced at world 2),
push (march reduces into the slab owner's buffer + k_adam_slab) with per-peer and with multicast parameter stores — with the
NCCL all-reduce trainer and the single-GPU trainer after 4 steps (replicas bit-equal, global loss, gathered optimiser
state, step() == step_host(), tv > 0), and (3) that a stalled peer ends in PlxError, not in a hang or a silently wrong grid.
Skipped on one-GPU boxes (the gloo world_size-2 tests in test_distributed_gloo.py cover the host-side sharding there;
bench.py repeats the replica / gradient self-check in its untimed region on every multi-GPU run).
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


def _run(case: str, timeout: int = 420) -> str:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multigpu_check.py"), case]
    out = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=timeout)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-4000:]
    return text


@pytest.mark.parametrize("case", ["pull", "pull_fused", "pull_mc", "push", "push_mc"])
def test_two_gpu_gradient_sum_and_peer_trainer(plx_lib, case):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    text = _run(case)
    assert "MULTIGPU_OK" in text and "PEER_OK" in text, text[-4000:]


def test_two_gpu_stalled_peer_raises_instead_of_corrupting(plx_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    text = _run("timeout")
    assert "TIMEOUT_OK" in text, text[-4000:]
