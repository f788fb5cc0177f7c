"""Multi-GPU parity (NCCL halo exchange, allreduce, host-array collectives): one process per GPU under torchrun.
Needs >= 2 GPUs; skipped on a single-GPU box (the world_size>1 HOST logic is covered on CPU by test_host_distributed)."""
import os
import subprocess
import sys

import pytest

from helpers import free_port

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_cases.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), script]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert f"multi-GPU parity ok on {world} ranks" in r.stdout
