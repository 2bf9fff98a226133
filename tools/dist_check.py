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
for (B, N, M, H, d, causal) in [(2, 200, 5000, 4, 64, False), (2, 512, 16384, 8, 128, False), (1, 96, 3000, 2, 128, True)]:
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
    bound = 1e-2 * ref.float().abs().max().item()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        print(f"shape {(B, N, M, H, d, causal)}: sharded vs single max err {err:.3e} (bound {bound:.3e}), identical on all ranks: {same}")
    ok = ok and err <= bound and same
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
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
