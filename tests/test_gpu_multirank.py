"""GPU, world size > 1 (skipped on a single-GPU box): on-hardware parity of the image-sharded path against the
single-GPU result over NCCL -- loss, gradients, integer tallies, VOC AP.  See tests/multirank_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = 29600 + world + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "failures 0" in res.stdout
