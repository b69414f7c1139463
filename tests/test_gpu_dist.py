"""Multi-GPU checks as collected tests: each one launches tests/dist_check.py under
`python -m torch.distributed.run` on the visible GPUs (one process per GPU, NCCL) and is skipped
when fewer than 2 GPUs are visible.  What dist_check.py asserts:

1. batch-sharded sampling over N ranks == per-seed single-process runs, bit-exact (no collective);
2. torch DDP x N at batch b == single process at batch N*b, gradient by gradient
   (ddpm_3d_ldm/train.py:232-233);
3. parallel.DistributedDataParallel (bucketed all-reduce overlapped with the backward launch
   list): eager, graph-capturing and replayed steps reproduce torch DDP's averaged gradients;
   no_sync() keeps gradients local and the next synchronised step reduces the accumulated ones.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4])
def test_dist_check_under_torchrun(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_check.py")]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "[dist_check] OK" in r.stdout, tail
