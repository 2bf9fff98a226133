"""world_size-2 `gloo` tests of the M-sharding protocol (perceiver_io_b200/dist.py) on CPU.

The device math is injected from the oracle (ShardKernels), so what is exercised here is the host-side
protocol: shard bounds, the MAX all-reduce, the rescale, the single packed SUM all-reduce and the final
normalisation — including ragged M, a shard that is fully padded for one batch row, and the causal mask
with global key offsets."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mha_oracle as O
from perceiver_io_b200.dist import ShardKernels, shard_bounds, sharded_attention


def test_shard_bounds_partition_the_key_axis():
    for m_total in (1, 127, 128, 129, 1000, 65536, 65537, 182528):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(m_total, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == m_total or any(e == m_total for _, e in cuts)
            covered = 0
            for b, e in cuts:
                assert b == covered or b == e == m_total
                assert e >= b
                assert b % 128 == 0 or b == e
                covered = max(covered, e)
            assert covered == m_total
            sizes = [e - b for b, e in cuts if e > b]
            assert max(sizes) - min(sizes) < 256  # tile counts differ by <= 1, last tile may be ragged


def _oracle_kernels(H):
    def partial(q, k, v, num_heads, scale, pad, causal, m_total, m_offset, out):
        qh = O.split_heads(q.double().expand(k.shape[0], -1, -1), num_heads)
        po, pm, pl = O.partial_state(qh, O.split_heads(k.double(), num_heads), O.split_heads(v.double(), num_heads),
                                     scale, pad, causal, m_total, m_offset)
        out[0].copy_(po)
        out[1].copy_(pm)
        out[2].copy_(pl)

    def rescale_(po, pm, pl, new_m):
        w = torch.exp2(pm - new_m)
        po.mul_(w[..., None])
        pl.mul_(w)
        pm.copy_(new_m)

    def finalize(po, pl, dtype):
        return O.merge_heads(po / pl[..., None]).to(dtype)

    return ShardKernels(partial=partial, rescale_=rescale_, finalize=finalize)


def _worker(rank, world, port, causal, result_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)          # identical "replicated" tensors on every rank
        B, H, N, M, d = 2, 2, 6, 300, 8
        q = torch.randn(B, N, H * d, generator=g) * 2
        k = torch.randn(B, M, H * d, generator=g)
        v = torch.randn(B, M, H * d, generator=g)
        pad = torch.zeros(B, M, dtype=torch.bool)
        pad[0, :256] = True          # rank 0's whole shard (and part of rank 1's) is padding for row 0
        pad[1, 290:] = True
        b, e = shard_bounds(M, world, rank)
        out = sharded_attention(q, k[:, b:e], v[:, b:e], H, d ** -0.5, M, b, pad[:, b:e], causal,
                                kernels=_oracle_kernels(H))
        ref = O.merge_heads(O.core_attention(O.split_heads(q.double(), H), O.split_heads(k.double(), H),
                                             O.split_heads(v.double(), H), d ** -0.5, pad, causal))
        result_queue.put((rank, (b, e), float((out.double() - ref).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("causal", [False, True])
def test_two_rank_gloo_merge_equals_unsharded(causal):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, causal, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[1] for r in results) == [(0, 256), (256, 300)]
    for rank, _, err in results:
        assert err < 1e-5, (rank, err)


def _grid_worker(rank, world, port, result_queue):
    """2-D rank grid (dist.plan_grid): B=2 on 4 ranks = 2 batch groups x 2 key shards.  Each batch group merges its own
    partial states over its own process sub-group; nothing crosses batch groups."""
    from perceiver_io_b200.dist import grid_position, m_shard_group, plan_grid

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B, H, N, M, d = 2, 2, 6, 300, 8
        q = torch.randn(B, N, H * d, generator=g) * 2
        k = torch.randn(B, M, H * d, generator=g)
        v = torch.randn(B, M, H * d, generator=g)
        pad = torch.zeros(B, M, dtype=torch.bool)
        pad[1, 290:] = True
        bg, mg = plan_grid(B, world)
        gb, gm = grid_position(rank, bg, mg)
        group = m_shard_group(bg, mg)
        rows = slice(gb * (B // bg), (gb + 1) * (B // bg))
        b, e = shard_bounds(M, mg, gm)
        out = sharded_attention(q[rows], k[rows, b:e], v[rows, b:e], H, d ** -0.5, M, b, pad[rows, b:e], False, group=group,
                                kernels=_oracle_kernels(H))
        ref = O.merge_heads(O.core_attention(O.split_heads(q.double(), H), O.split_heads(k.double(), H),
                                             O.split_heads(v.double(), H), d ** -0.5, pad, False))[rows]
        result_queue.put((rank, (bg, mg, gb, gm), dist.get_world_size(group), float((out.double() - ref).abs().max())))
    finally:
        dist.destroy_process_group()


def test_four_rank_gloo_batch_by_key_grid():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, 4, port, queue)) for r in range(4)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[1] for r in results) == [(2, 2, 0, 0), (2, 2, 0, 1), (2, 2, 1, 0), (2, 2, 1, 1)]
    for rank, _, group_size, err in results:
        assert group_size == 2 and err < 1e-5, (rank, group_size, err)
