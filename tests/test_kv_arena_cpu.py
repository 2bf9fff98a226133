"""CPU tests of the arena bookkeeping behind ops.kv_append (host logic only: the launch is replaced by a torch
stand-in).  What must hold is the reference's functional contract (modules.py:117-121 is a torch.cat): the result
equals the concatenation, and nothing a caller still holds ever changes — also when a cache is continued twice,
truncated to a sliding window (core/huggingface.py:146-156) or batch-reordered (:140-144)."""
import pytest
import torch

from perceiver_io_b200 import ops


@pytest.fixture
def cpu_launch(monkeypatch):
    calls = {"in_place": 0, "copied": 0}

    def fake_launch(kc, vc, kn, vn, kd, vd, k_in_place, v_in_place):
        L = kc.shape[1]
        for cache, new, dst, in_place in ((kc, kn, kd, k_in_place), (vc, vn, vd, v_in_place)):
            if in_place:
                calls["in_place"] += 1
            elif L:
                dst[:, :L] = cache
                calls["copied"] += L
            dst[:, L:] = new

    monkeypatch.setattr(ops, "_launch_kv_append", fake_launch)
    monkeypatch.setattr(ops, "_require_cuda", lambda *a: None)
    return calls


def test_decode_loop_appends_in_place_and_matches_cat(cpu_launch):
    B, Ck, Cv = 2, 16, 24
    k, v = torch.zeros(B, 0, Ck), torch.zeros(B, 0, Cv)
    ks, vs = [], []
    for _ in range(300):
        kn, vn = torch.randn(B, 1, Ck), torch.randn(B, 1, Cv)
        ks.append(kn)
        vs.append(vn)
        k, v = ops.kv_append(k, v, kn, vn)
        assert torch.equal(k, torch.cat(ks, 1)) and torch.equal(v, torch.cat(vs, 1))
    # amortised: far fewer rows copied than the 300*299/2 of a concat per step, most appends in place
    assert cpu_launch["copied"] < 2 * 4 * 300 and cpu_launch["in_place"] > 2 * 280


def test_two_continuations_of_one_cache_stay_independent(cpu_launch):
    B, C = 2, 8
    k, v = ops.kv_append(torch.zeros(B, 0, C), torch.zeros(B, 0, C), torch.randn(B, 5, C), torch.randn(B, 5, C))
    snap = k.clone()
    a, _ = ops.kv_append(k, v, torch.full((B, 1, C), 1.0), torch.full((B, 1, C), 1.0))
    b, _ = ops.kv_append(k, v, torch.full((B, 1, C), 2.0), torch.full((B, 1, C), 2.0))
    assert torch.equal(k, snap)
    assert torch.equal(a[:, :5], snap) and torch.equal(b[:, :5], snap)
    assert (a[:, 5] == 1).all() and (b[:, 5] == 2).all()
    # continuing the OLDER branch again must not disturb the newer one either
    c, _ = ops.kv_append(a, a, torch.full((B, 2, C), 3.0), torch.full((B, 2, C), 3.0))
    assert (b[:, 5] == 2).all() and (c[:, 5] == 1).all() and (c[:, 6:] == 3).all()


def test_sliding_window_truncation_and_reorder(cpu_launch):
    B, C, W = 3, 8, 20
    full = torch.randn(B, W, C)
    k, v = ops.kv_append(torch.zeros(B, 0, C), torch.zeros(B, 0, C), full, full)
    for step in range(200):
        kn = torch.randn(B, 1, C)
        full = torch.cat([full, kn], 1)
        k2, v2 = ops.kv_append(k, v, kn, kn)
        assert torch.equal(k2, full[:, -(W + 1):]) and torch.equal(v2, k2)
        k, v = k2[:, -W:], v2[:, -W:]            # what _truncate_*_past_key_values do
        if step % 37 == 0:                        # what _reorder_cache does
            idx = torch.randperm(B)
            k, v, full = k.index_select(0, idx), v.index_select(0, idx), full.index_select(0, idx)


def test_foreign_views_are_never_written(cpu_launch):
    """A cache that merely LOOKS like an arena row range (a slice of some user tensor with slack) is copied."""
    B, C = 2, 8
    user = torch.randn(B, 100, C)
    snap = user.clone()
    k, _ = ops.kv_append(user[:, :10], user[:, :10], torch.ones(B, 1, C), torch.ones(B, 1, C))
    assert torch.equal(user, snap) and torch.equal(k[:, :10], snap[:, :10])


def test_arena_can_be_disabled(cpu_launch, monkeypatch):
    monkeypatch.setitem(ops.kv_arena_config, "enabled", False)
    B, C = 1, 8
    k, v = ops.kv_append(torch.zeros(B, 0, C), torch.zeros(B, 0, C), torch.randn(B, 3, C), torch.randn(B, 3, C))
    k2, _ = ops.kv_append(k, v, torch.randn(B, 1, C), torch.randn(B, 1, C))
    assert k2.is_contiguous() and k2._base is None and k2.shape == (B, 4, C)
