"""Multi-GPU path (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on one GPU).
The sharded schedule + multi-GPU fusion must reproduce the single-GPU (Jacobi-ordered) result bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_and_fusion_match_single_gpu(world, apde_lib):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29710 + world), os.path.join(ROOT, "tools", "dist_fusion_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "DIST_FUSION_CHECK PASS" in r.stdout
