"""NCCL tests over all GPUs of the box (skipped on a single-GPU box): data-parallel gradients and tile-sharded
inference against the single-process results -- see tests/multi_gpu_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_nccl_gradients_and_tile_sharded_assignments_match_single_process():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 8)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=580, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MULTI_GPU_OK {n}" in r.stdout
