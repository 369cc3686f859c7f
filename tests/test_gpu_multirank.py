"""GPU, N > 1: halo exchange and cross-rank image reduction over NCCL, one process per GPU.

Skipped on single-GPU boxes (the world_size-2 gloo test covers the host logic there)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,dims", [(2, "1x1x2"), (2, "1x2x1"), (4, "1x2x2"), (4, "1x1x4"), (8, "1x4x2")])
def test_multi_gpu_image_equals_single_rank_oracle(world, dims):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() + world * 7 + len(dims)) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py"), dims]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "ok" in proc.stdout
