"""CPU test of the multi-GPU benchmark scene (bench.py: scene_layout / build_scene): ONE scene whose nearest-camera source lists
cross the view blocks the ranks own, at every GPU count the driver runs -- so every geometric pass of `bench.py --gpus N` reads
depth maps computed on other GPUs (round 1 benchmarked N independent replicas)."""
import numpy as np
import pytest

from helpers import ROOT  # noqa: F401  (puts the repo root on sys.path)


def _block(V, world, rank):  # apde_view_block, apde_context.h
    base, extra = divmod(V, world)
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_source_lists_cross_rank_blocks(world):
    from apde_mvs_b200.scene import make_office_scene
    per_arc, src = 11, 10
    scene = make_office_scene(64, 48, per_arc, src, seed=2, weak=0.0, rings=world)
    V = per_arc * world
    assert len(scene.images) == V and all(len(p) == src for p in scene.pairs)
    total = 0
    for rank in range(world):
        first, count = _block(V, world, rank)
        assert (first, count) == (rank * per_arc, per_arc)  # rank r owns arc r
        cross = sum(1 for v in range(first, first + count) for s in scene.pairs[v] if not (first <= s < first + count))
        assert cross >= per_arc, "rank %d of %d reads only %d maps of other ranks" % (rank, world, cross)
        total += cross
    assert total >= 0.5 * V * src  # most source reads cross rank boundaries (measured on the B200 runs: 126/220, 352/440, 712/880)
    # and no view lists itself or a view outside the scene
    for v, p in enumerate(scene.pairs):
        assert v not in p and min(p) >= 0 and max(p) < V and len(set(p)) == len(p)
