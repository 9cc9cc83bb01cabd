"""One process per GPU over a single scene: reference views are sharded in contiguous blocks, the owned
depth maps are all-gathered between passes, and before fusion the normal / weak / confidence maps follow.  WeakVisFilter
(APD.cpp:962-1049, the reference's thread-pool work items) is sharded by reference view; the greedy fusion itself is
ordered over views (APD.cpp:1149,1176,1209) and runs on rank 0 over the gathered maps and skip masks.

Every collective runs IN PLACE on the library's device pools (apde_map_pool), wrapped as torch tensors through
__cuda_array_interface__: NCCL moves the rows over NVLink, nothing is staged through the host."""
import numpy as np

from .binding import POOL
from .sharding import exchange_rows, shard

_DTYPE = {POOL.DEPTH: ("<f4", 4), POOL.NORMAL: ("<f4", 4), POOL.WEAK: ("|u1", 1), POOL.CONFIDENCE: ("|u1", 1), POOL.SKIP: ("|u1", 1)}


class DistributedScene:
    def __init__(self, ctx, dist, device):
        self.ctx, self.dist, self.device = ctx, dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.first, self.count = shard(ctx.V, self.world, self.rank)
        self._tensors = {}

    def pool_tensor(self, which):
        """the [V, elements per view] device pool of one field as a torch tensor sharing the library's memory"""
        import torch
        ptr, total, per = self.ctx.map_pool(which)
        key = (which, ptr, per)
        if key not in self._tensors:
            typestr, size = _DTYPE[which]

            class _Wrap:
                __cuda_array_interface__ = {"shape": (self.ctx.V, per // size), "typestr": typestr, "data": (ptr, False), "version": 2}
            self._tensors[key] = torch.as_tensor(_Wrap(), device=self.device)
        return self._tensors[key]

    def exchange(self, which):
        import torch
        t = self.pool_tensor(which)
        exchange_rows(self.dist, t, self.ctx.V, self.world)
        if t.is_cuda:
            torch.cuda.synchronize()

    def run_schedule(self, sched, timing=None):
        """all passes; this rank runs its shard of reference views (Jacobi ordering) and the depth maps are exchanged"""
        from .binding import Timing
        sched.jacobi = 1 if self.world > 1 else sched.jacobi
        sched.first_view, sched.num_views_local = self.first, self.count
        t = timing if timing is not None else Timing()
        for p in range(self.ctx.num_passes(sched)):
            self.ctx.run_schedule_pass(sched, p, t)
            if self.world > 1:
                self.exchange(POOL.DEPTH)
        return t

    def gather_maps(self):
        """normal / weak / confidence maps of every view on every rank (depth is already replicated)"""
        d = self.ctx.view_dims(self.first) if self.count else (0, 0)
        dims = [None] * self.world
        self.dist.all_gather_object(dims, d)
        w, h = max(x[0] for x in dims), max(x[1] for x in dims)
        if self.world > 1:
            for which in (POOL.NORMAL, POOL.WEAK, POOL.CONFIDENCE):
                self.exchange(which)
            self.ctx.views_mark_maps(w, h)
        return w, h

    def fuse(self, weak_filter=True, variant=0):
        """returns (xyz, bgr) on rank 0 and (None, None) elsewhere"""
        self.gather_maps()
        if weak_filter:
            self.ctx.weak_vis_filter_range(self.first, self.count)
            if self.world > 1:
                self.exchange(POOL.SKIP)
        self.dist.barrier()
        if self.rank != 0:
            return None, None
        return self.ctx.fuse(2 if weak_filter else 0, variant=variant)
