"""Small launch mix for compute-sanitizer (memcheck): every kernel family once on ragged shapes, so that an
out-of-bounds or misaligned access in a tail tile would be reported.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)


def qkv(B, N, M, H, dqk, dv):
    return (torch.randn(B, N, H * dqk, generator=g).bfloat16().cuda(), torch.randn(B, M, H * dqk, generator=g).bfloat16().cuda(),
            torch.randn(B, M, H * dv, generator=g).bfloat16().cuda())


CASES = [("tcgen05 128/128 ragged + masks", (2, 200, 333, 2, 128, 128), "tcgen05", True),
         ("tcgen05 64/64 split units", (1, 130, 1500, 2, 64, 64), "tcgen05", False),
         ("tcgen05 wide dv 192", (1, 130, 300, 2, 64, 192), "tcgen05", False),
         ("tcgen05 big head 328", (1, 40, 200, 1, 328, 328), "tcgen05", False),
         ("tcgen05 decode N=1", (2, 1, 777, 4, 96, 96), "tcgen05", True),
         ("simt odd dims", (2, 33, 100, 2, 24, 40), "simt", True)]
for name, (B, N, M, H, dqk, dv), impl, masks in CASES:
    q, k, v = qkv(B, N, M, H, dqk, dv)
    pad = None
    if masks:
        pad = torch.zeros(B, M, dtype=torch.bool)
        pad[0, :37] = True
        pad = pad.cuda()
    out = ops.attention(q, k, v, H, dqk ** -0.5, pad_mask=pad, causal=masks, impl=impl)
    parts = [ops.attention_partial(q, k[:, a:b], v[:, a:b], H, dqk ** -0.5, m_total=M, m_offset=a, impl=impl)
             for a, b in ((0, M // 2), (M // 2, M))]
    merged = ops.combine_partials(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                  torch.stack([p[2] for p in parts]))
    torch.cuda.synchronize()
    print("ok", name, float(out.float().abs().mean()), float(merged.float().abs().mean()), flush=True)
kc, vc = torch.zeros(2, 0, 64, device="cuda").bfloat16(), torch.zeros(2, 0, 64, device="cuda").bfloat16()
for n in (70, 1, 1, 1):
    kc, vc = ops.kv_append(kc, vc, torch.randn(2, n, 64, generator=g).bfloat16().cuda(), torch.randn(2, n, 64, generator=g).bfloat16().cuda())
ang = torch.randn(1, 73, 32, generator=g).cuda()
r = ops.rotary(kc, 2, ang, True)
torch.cuda.synchronize()
print("ok aux", tuple(kc.shape), tuple(r.shape), flush=True)
print("SANITIZE_TARGET_DONE")
