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


def test_more_ranks_than_views(apde_lib):
    """3 views over 4 GPUs: one rank holds no view, takes part in every exchange round and in the collective fusion, and the job
    still equals the single-GPU Jacobi run bit for bit"""
    if _gpus() < 4:
        pytest.skip("needs 4 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", "29719", os.path.join(ROOT, "tools", "dist_fusion_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, APDE_CHECK_VIEWS="3"))
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "DIST_FUSION_CHECK PASS" in r.stdout


def _read_all(folder, V):
    out = {}
    for v in range(V):
        for name in ("depths", "normals", "weak", "confidence"):
            out[(v, name)] = open(os.path.join(folder, "APD", "%08d" % v, name + ".bin"), "rb").read()
        out[(v, "skip")] = open(os.path.join(folder, "APD", "%08d" % v, "skip.png"), "rb").read()
    out["ply"] = open(os.path.join(folder, "APD", "APD.ply"), "rb").read()
    return out


@pytest.mark.parametrize("world", [2, 4])
def test_apd_cli_job_equals_single_gpu_jacobi(world, tmp_path, apde_lib):
    """`apd --gpus N` (C++ only: host threads + NCCL inside libapde, no Python on the path) on a dense folder writes byte for
    byte the files of `apd --view_order jacobi` on one GPU: maps of all views, skip masks, fused cloud"""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import shutil
    sys.path.insert(0, ROOT)
    from apde_mvs_b200 import build as b
    from apde_mvs_b200.scene import make_office_scene
    apd = b.build_host()
    V = 11  # ragged blocks; 10 source views: the strong propagation needs its opt-in to > 48 KB of shared memory on EVERY device
    scene = make_office_scene(480, 360, num_views=V, num_src=10, seed=6, weak=0.2, with_color=True)
    one, many = str(tmp_path / "one"), str(tmp_path / "many")
    scene.write_dense_folder(one)
    shutil.copytree(one, many)
    env = dict(os.environ, APDE_NO_SHOW="1")
    r1 = subprocess.run([apd, "-d", one, "--view_order", "jacobi", "--dataset", "General"], capture_output=True, text=True, env=env, timeout=900)
    assert r1.returncode == 0, r1.stdout[-2000:]
    rn = subprocess.run([apd, "-d", many, "--gpus", str(world), "--dataset", "General"], capture_output=True, text=True, env=env, timeout=900)
    print(rn.stdout[-1500:])
    assert rn.returncode == 0, rn.stdout[-2000:] + rn.stderr[-2000:]
    assert "Depth-map exchange:" in rn.stdout
    a, m = _read_all(one, V), _read_all(many, V)
    diff = [k for k in a if a[k] != m[k]]
    assert not diff, diff
