"""torchrun --nproc-per-node N tools/dist_fusion_check.py : a multi-GPU job (views dealt out over the ranks, depth maps exchanged
over NCCL inside libapde, collective fusion) must give, on rank 0, exactly the maps and the cloud a single GPU computes with
the same (Jacobi) view ordering.  torch.distributed (gloo) only carries the 128-byte NCCL id and the verdict."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from apde_mvs_b200.binding import Context, default_schedule
from apde_mvs_b200.distributed import Job
from apde_mvs_b200.scene import make_office_scene

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
W, H, V, N = 480, 360, 7, 4  # 7 views over 2 or 4 ranks: ragged shards
if os.environ.get("APDE_CHECK_VIEWS"):  # e.g. 3 views over 4 ranks: one rank holds no view and only receives
    V = int(os.environ["APDE_CHECK_VIEWS"])
    N = min(N, V - 1)
scene = make_office_scene(W, H, num_views=V, num_src=N, seed=6, weak=0.2, with_color=True)
ctx = Context(local)
ctx.load_scene(scene)
sched = default_schedule()
sched.rounds, sched.seed = 2, 13
ds = Job(ctx, dist)
t0 = time.time()
tm = ds.run_schedule(sched)
print("rank %d: block [%d, %d), exposed exchange %.2f ms, %.1f MB received" % (rank, ds.first, ds.first + ds.count, tm.exchange_ms, tm.exchange_bytes / 1e6))
out = {}
for variant in (0, 1, 2):
    xyz, bgr = ds.fuse(True, variant=variant)
    out[variant] = (xyz, bgr)
t1 = time.time()
ok = True
if rank == 0:
    ref = Context(local)
    ref.load_scene(scene)
    s2 = default_schedule()
    s2.rounds, s2.seed, s2.jacobi = 2, 13, 1
    ref.run_schedule(s2)
    for v in range(V):
        a, b = ctx.view_download(v), ref.view_download(v)
        for k in range(4):
            if not np.array_equal(a[k], b[k]):
                ok = False
                print("view %d map %d differs between the %d-GPU run and the single-GPU Jacobi run" % (v, k, world))
    for variant in (0, 1, 2):
        xyz_r, bgr_r = ref.fuse(True, variant=variant)
        xyz, bgr = out[variant]
        same = len(xyz) == len(xyz_r) and np.array_equal(xyz, xyz_r) and np.array_equal(bgr, bgr_r)
        print("variant %d: %d-GPU fusion %d points, single GPU %d points, identical %s" % (variant, world, len(xyz), len(xyz_r), same))
        ok &= bool(same) and len(xyz) > 1000
    print("DIST_FUSION_CHECK %s world=%d %.1fs" % ("PASS" if ok else "FAIL", world, t1 - t0))
flag = torch.tensor([1 if ok else 0])
dist.broadcast(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
