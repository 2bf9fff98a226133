"""Multi-GPU check (run under torchrun, one rank per GPU): the M-sharded cross-attention equals the single-GPU
result and the oracle.  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import perceiver_io_b200 as P  # noqa: E402
from perceiver_io_b200 import ops  # noqa: E402
from perceiver_io_b200.dist import cross_attention_sharded, shard_bounds, sharded_attention  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (B, N, M, H, d, causal) in [(2, 200, 5000, 4, 64, False), (2, 512, 16384, 8, 128, False), (1, 96, 3000, 2, 128, True),
                                (2, 300, 8192, 2, 128, True), (8, 512, 8192, 8, 128, False), (1, 200, 4096, 4, 64, False)]:
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B, N, H * d, generator=g) * 2).bfloat16().to(dev)
    k = torch.randn(B, M, H * d, generator=g).bfloat16().to(dev)
    v = torch.randn(B, M, H * d, generator=g).bfloat16().to(dev)
    pad = torch.zeros(B, M, dtype=torch.bool)
    pad[0, : M // 3] = True
    pad = pad.to(dev)
    m0, m1 = shard_bounds(M, world, rank)
    out = sharded_attention(q, k[:, m0:m1], v[:, m0:m1], H, d ** -0.5, M, m0, pad[:, m0:m1], causal, merge="peer")
    out_nccl = sharded_attention(q, k[:, m0:m1], v[:, m0:m1], H, d ** -0.5, M, m0, pad[:, m0:m1], causal, merge="nccl")
    ref = ops.attention(q, k, v, H, d ** -0.5, pad_mask=pad, causal=causal)
    err = max((out.float() - ref.float()).abs().max().item(), (out_nccl.float() - ref.float()).abs().max().item())
    if M % (128 * world) == 0:
        # fused kernel-tail merge (one launch per rank): several calls in a row exercise the epoch flags and buffer reuse
        for it in range(4):
            qq = q if it % 2 == 0 else (q * 0.5).to(q.dtype)
            fused = sharded_attention(qq, k[:, m0:m1], v[:, m0:m1], H, d ** -0.5, M, m0, pad[:, m0:m1], causal, merge="fused")
            ref_it = ref if it % 2 == 0 else ops.attention(qq, k, v, H, d ** -0.5, pad_mask=pad, causal=causal)
            e = (fused.float() - ref_it.float()).abs().max().item()
            err = max(err, e)
            peer_it = sharded_attention(qq, k[:, m0:m1], v[:, m0:m1], H, d ** -0.5, M, m0, pad[:, m0:m1], causal, merge="peer")
            if not torch.equal(fused, peer_it):   # same arithmetic, same order over ranks: bit-identical
                err = max(err, float("inf") if (fused.float() - peer_it.float()).abs().max().item() > 1e-2 else err)
                if rank == 0:
                    print(f"  note: fused != peer bitwise at call {it}: max diff {(fused.float() - peer_it.float()).abs().max().item():.3e}")
    bound = 1e-2 * ref.float().abs().max().item()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        print(f"shape {(B, N, M, H, d, causal)}: sharded vs single max err {err:.3e} (bound {bound:.3e}), identical on all ranks: {same}")
    ok = ok and err <= bound and same
# decode step against a sharded cache: one query row, partial states from the streaming decode kernel
for (B, N, M, H, d) in [(4, 1, 16384, 8, 128), (2, 2, 8192, 4, 64)]:
    g = torch.Generator().manual_seed(9)
    q = torch.randn(B, N, H * d, generator=g).bfloat16().to(dev)
    k = torch.randn(B, M, H * d, generator=g).bfloat16().to(dev)
    v = torch.randn(B, M, H * d, generator=g).bfloat16().to(dev)
    m0, m1 = shard_bounds(M, world, rank)
    out = sharded_attention(q, k[:, m0:m1], v[:, m0:m1], H, d ** -0.5, M, m0, causal=True)   # auto -> peer for N <= 4
    ref = ops.attention(q, k, v, H, d ** -0.5, causal=True)
    err = (out.float() - ref.float()).abs().max().item()
    bound = 1e-2 * ref.float().abs().max().item()
    if rank == 0:
        print(f"sharded decode step {(B, N, M, H, d)}: max err {err:.3e} (bound {bound:.3e})")
    ok = ok and err <= bound
# module-level entry point
torch.manual_seed(0)
layer = P.CrossAttention(8, 256, 256).to(dev).bfloat16().eval()
for prm in layer.parameters():
    dist.broadcast(prm.data, src=0)
xq = torch.randn(1, 128, 256, device=dev).bfloat16()
xkv = torch.randn(2, 4096, 256, device=dev).bfloat16()
dist.broadcast(xq, src=0)
dist.broadcast(xkv, src=0)
m0, m1 = shard_bounds(4096, world, rank)
with torch.no_grad():
    a = cross_attention_sharded(layer, xq, xkv[:, m0:m1], 4096, m0).last_hidden_state
    b = layer(xq, xkv).last_hidden_state
err = (a.float() - b.float()).abs().max().item()
if rank == 0:
    print(f"cross_attention_sharded vs CrossAttention.forward: max err {err:.3e}")
ok = ok and err <= 2e-2 * b.float().abs().max().item()
# timing at the north-star shape: fused (one launch) vs peer (3 launches + 2 host-launched barriers)
if os.environ.get("PCV_DIST_TIMING", "1") == "1":
    B, N, M, H, d = 8, 512, 65536, 8, 128
    m0, m1 = shard_bounds(M, world, rank)
    torch.manual_seed(rank)
    q = torch.randn(B, N, H * d, device=dev).bfloat16()
    dist.broadcast(q, src=0)
    k = torch.randn(B, m1 - m0, H * d, device=dev).bfloat16()
    v = torch.randn(B, m1 - m0, H * d, device=dev).bfloat16()
    for merge in ("fused", "peer"):
        for _ in range(5):
            sharded_attention(q, k, v, H, d ** -0.5, M, m0, merge=merge, copy_out=False)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            sharded_attention(q, k, v, H, d ** -0.5, M, m0, merge=merge, copy_out=False)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 30], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0 and merge == "fused":
            from perceiver_io_b200.dist import PeerMerger
            for key, pmx in PeerMerger._cache.items():
                if key[-1] is not None and key[-1][0] == "fused" and pmx.B == B:
                    st = pmx.flags[24:30].tolist()
                    print(f"  fused tail phase clock of CTA 0 (us since tail entry): grid arrival {st[0] / 1e3:.1f}, local split merge "
                          f"{st[1] / 1e3:.1f}, peers ready {st[2] / 1e3:.1f}, rows merged+pushed {st[3] / 1e3:.1f}, grid arrival "
                          f"{st[4] / 1e3:.1f}, all peers done {st[5] / 1e3:.1f}")
        if rank == 0:
            print(f"north-star shape on {world} GPUs, merge={merge}: {t.item():.4f} ms/step = {4.0 * B * N * M * H * d / t.item() / 1e9:.0f} TFLOP/s aggregate")
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
