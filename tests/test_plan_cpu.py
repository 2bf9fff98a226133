"""CPU property tests of the stream-K work plan of the tcgen05 kernels (host logic in csrc/pcv_attn_tc.cu,
dumped through the host-only C-ABI entry point pcv_debug_plan; no GPU involved).  The plan decides which CTA
multiplies which (batch, head, query block) with which key tiles and which partial slots the combine kernel
merges — a hole or an overlap here is a silent wrong answer, so the invariants are checked over random shapes."""
import collections

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from perceiver_io_b200 import _lib

TILE = 128


def _check(B, H, N, M, workers, rpu):
    counts, segs = _lib.debug_plan(B, H, N, M, workers=workers, rows_per_unit=rpu)
    T = (M + TILE - 1) // TILE
    QB = (N + rpu - 1) // rpu
    assert counts["segments"] == len(segs) > 0
    assert 0 < counts["ctas"] <= workers
    per_unit = collections.defaultdict(list)
    per_cta = collections.defaultdict(list)
    for cta, b, h, q0, ntile, t0, t1, slot in segs:
        assert 0 <= cta < counts["ctas"] and 0 <= b < B and 0 <= h < H
        assert q0 % rpu == 0 and 0 <= q0 < N
        assert 0 <= t0 < t1 <= T
        assert ntile == (2 if (rpu > TILE and N - q0 > TILE) else 1)
        assert (slot == -1) == (t0 == 0 and t1 == T)
        per_unit[(b, h, q0)].append((t0, t1, slot))
        per_cta[cta].append((b * H + h, q0, t0, t1))
    # every (b, h, query block) sees every key tile exactly once
    assert len(per_unit) == B * H * QB
    slots_seen = []
    split_units = 0
    for key, pieces in per_unit.items():
        pieces.sort()
        assert pieces[0][0] == 0 and pieces[-1][1] == T, (key, pieces)
        for (a0, a1, _), (b0, b1, _) in zip(pieces, pieces[1:]):
            assert a1 == b0, (key, pieces)
        s = sorted(p[2] for p in pieces)
        if len(pieces) == 1:
            assert s == [-1]
        else:
            split_units += 1
            assert s[0] >= 0 and s == list(range(s[0], s[0] + len(s))), (key, s)  # contiguous for the combine kernel
            slots_seen += s
    assert split_units == counts["units"]
    assert sorted(slots_seen) == list(range(counts["slots"]))
    # every CTA has work, and the load is balanced to within one tile (split mode) / one unit (whole-unit mode)
    assert set(per_cta) == set(range(counts["ctas"]))
    tiles = [sum(t1 - t0 for _, _, t0, t1 in v) for v in per_cta.values()]
    if QB * 2 <= workers:
        assert max(tiles) - min(tiles) <= 1, (min(tiles), max(tiles))
        # the QB members of a group walk the same (b*h, key range) sequence -> they share their K/V stream in L2
        for g in range(counts["ctas"] // QB):
            walks = [[(bh, t0, t1) for bh, _, t0, t1 in per_cta[g * QB + r]] for r in range(QB)]
            assert all(w == walks[0] for w in walks)
    else:
        assert max(tiles) - min(tiles) <= T


@pytest.mark.parametrize("shape", [
    (8, 8, 512, 65536, 148, 256),    # north star
    (8, 8, 512, 65536 // 8, 148, 256),   # one of 8 M-shards
    (1, 1, 2048, 182528, 148, 128),  # optical-flow encoder (big-head kernel: one query tile per unit)
    (1, 1, 182528, 2048, 148, 128),  # optical-flow decoder: more query blocks than CTAs
    (8, 8, 1, 16384, 148, 256),      # decode step
    (2, 2, 300, 333, 4, 256),
    (1, 1, 1, 1, 148, 256),
    (3, 5, 1000, 77, 7, 512),        # CTA-pair units on an odd worker count
], ids=lambda s: "x".join(map(str, s)))
def test_plan_invariants_named_shapes(shape):
    _check(*shape)


@settings(max_examples=300, deadline=None)
@given(B=st.integers(1, 9), H=st.integers(1, 9), N=st.integers(1, 3000), M=st.integers(1, 70000),
       workers=st.sampled_from([1, 2, 3, 4, 7, 16, 74, 148, 160]), rpu=st.sampled_from([128, 256, 512]))
def test_plan_invariants_random_shapes(B, H, N, M, workers, rpu):
    _check(B, H, N, M, workers, rpu)


def test_debug_plan_rejects_bad_arguments():
    with pytest.raises(_lib.PcvError):
        _lib.debug_plan(0, 1, 1, 1)
    with pytest.raises(_lib.PcvError):
        _lib.debug_plan(1, 1, 1, 1, rows_per_unit=100)
