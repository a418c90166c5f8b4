"""NCCL check of the row-sharded gallery (`distributed.ShardedGallery`): self-launches `torch.distributed.run` with one
rank per visible GPU (2, 4 or 8) and requires search + ranks to equal the single-GPU result bit for bit
(tests/run_multigpu_check.py).  Skipped on a single-GPU box; the collective logic itself is covered on CPUs with gloo
(tests/test_distributed_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_row_sharded_search_and_ranks_equal_single_gpu():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "run_multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, f"{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
