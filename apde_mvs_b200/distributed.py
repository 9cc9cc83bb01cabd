"""One process per GPU over a single scene.  The data path is in the library: libapde deals the reference views
out in contiguous blocks, broadcasts every finished depth map over NCCL while the next view is computed, and gathers the
normal / weak / confidence / skip maps before fusion (apde_comm_*, apde_exchange, apde_fuse_collective; csrc/apde_comm.cu).

What is left for the launcher -- torchrun here, any MPI-like launcher elsewhere -- is the rendezvous: rank 0 makes the
128-byte NCCL id and hands it to the other ranks.  `Job` does that through a torch.distributed process group (gloo is enough:
the group carries 128 bytes and the timing reductions of bench.py, never a map)."""
from .binding import COMM_ID_BYTES


def share_comm_id(ctx_cls, dist):
    """rank 0 creates the id of the job; every rank returns the same 128 bytes"""
    box = [ctx_cls.comm_create_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm_id = box[0]
    assert isinstance(comm_id, (bytes, bytearray)) and len(comm_id) == COMM_ID_BYTES
    return bytes(comm_id)


class Job:
    """a context that is one rank of a multi-GPU job: the schedule and the fusion become collective calls"""

    def __init__(self, ctx, dist):
        self.ctx, self.dist = ctx, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 1:
            ctx.comm_init(share_comm_id(type(ctx), dist), self.rank, self.world)
        _, _, self.first, self.count = ctx.comm_info()

    def run_schedule(self, sched, timing=None):
        """all passes: this rank runs its block of reference views, the library exchanges the depth maps (Jacobi order)"""
        from .binding import Timing
        t = timing if timing is not None else Timing()
        for p in range(self.ctx.num_passes(sched)):
            self.ctx.run_schedule_pass(sched, p, t)
        return t

    def fuse(self, weak_filter=True, variant=0):
        """returns (xyz, bgr) on rank 0 and (None, None) elsewhere"""
        return self.ctx.fuse_collective(weak_filter, variant=variant)
