"""Multi-GPU host logic: reference views are independent inside a pass once the previous pass's depth maps are fixed
(Jacobi ordering), so views are sharded in contiguous blocks over the ranks and the owned depth maps
are all-gathered between passes.  The collective runs on the library's replicated depth pool in place."""


def shard(num_views, world, rank):
    """contiguous block [first, first + count) of reference views owned by `rank`; blocks differ by at most one view"""
    base, rem = divmod(num_views, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def exchange_rows(dist, pool, num_views, world):
    """pool: [V, X] tensor (any dtype) replicated on every rank, each rank has fresh rows for its own shard only.
    Equal shards use one in-place all_gather_into_tensor; ragged shards fall back to per-rank broadcasts."""
    base, rem = divmod(num_views, world)
    rank = dist.get_rank()
    if rem == 0:
        first, count = shard(num_views, world, rank)
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(pool, pool[first:first + count])
        else:  # gloo: no in-place aliasing
            dist.all_gather_into_tensor(pool, pool[first:first + count].clone())
        return
    for r in range(world):
        first, count = shard(num_views, world, r)
        if count:
            dist.broadcast(pool[first:first + count], src=r)


def exchange_depth_maps(dist, pool, num_views, world):
    """the per-pass exchange of the replicated depth pool (kept under its first name)"""
    exchange_rows(dist, pool, num_views, world)
