"""How a multi-GPU job deals the reference views out: contiguous blocks, the first V % world ranks hold one
view more.  The library does the same in apde_comm_block_of (csrc/apde_context.h: apde_view_block); this mirror lets host code
plan (which rank writes which view's files) without a context."""


def shard(num_views, world, rank):
    """contiguous block [first, first + count) of reference views owned by `rank`; blocks differ by at most one view"""
    base, rem = divmod(num_views, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)
