"""CPU: the rank grid of the multi-GPU decomposition (dist.plan_grid): batch axis first, M shards for the rest."""
import pytest

from perceiver_io_b200.dist import grid_position, plan_grid, shard_bounds


@pytest.mark.parametrize("batch,world,expect", [(8, 1, (1, 1)), (8, 2, (2, 1)), (8, 4, (4, 1)), (8, 8, (8, 1)),
                                                (1, 8, (1, 8)), (2, 8, (2, 4)), (6, 8, (2, 4)), (3, 4, (1, 4)),
                                                (12, 8, (4, 2)), (5, 5, (5, 1))])
def test_plan_grid(batch, world, expect):
    bg, mg = plan_grid(batch, world)
    assert (bg, mg) == expect
    assert bg * mg == world and batch % bg == 0
    # every (batch row, key) pair is owned by exactly one rank
    M = 1000
    seen = set()
    for r in range(world):
        gb, gm = grid_position(r, bg, mg)
        rows = range(gb * (batch // bg), (gb + 1) * (batch // bg))
        m0, m1 = shard_bounds(M, mg, gm)
        for b in rows:
            for blk in range(m0, m1, 50):
                assert (b, blk) not in seen
                seen.add((b, blk))
    assert len({b for b, _ in seen}) == batch


def test_plan_grid_rejects_bad_input():
    with pytest.raises(ValueError):
        plan_grid(0, 4)
    with pytest.raises(ValueError):
        grid_position(8, 2, 4)
