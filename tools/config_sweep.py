"""Core-attention timing over the shapes of BASELINE.json's configs (B200): which kernel family serves each,
time, TFLOP/s and algorithmic GB/s.  Not a bench line — a coverage table for profiles/."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import _lib, ops  # noqa: E402

ONLY = os.environ.get("PCV_SWEEP_ONLY")  # substring filter
CASES = [
    # name, B, N, M, H, dqk, dv, causal
    ("mnist enc x-attn (N=32,M=784,dh=131)", 8, 32, 784, 1, 131, 131, False),
    ("mnist self-attn (N=32,dh=16)", 8, 32, 32, 8, 16, 16, False),
    ("mlm enc x-attn (N=256,M=2048,32/160)", 8, 256, 2048, 8, 32, 160, False),
    ("mlm self-attn (N=256,32/160)", 8, 256, 256, 8, 32, 160, False),
    ("mlm dec x-attn (O=2048,N=256,32/96)", 8, 2048, 256, 8, 32, 96, False),
    ("optical-flow enc x-attn (N=2048,M=182528,dh=322)", 1, 2048, 182528, 1, 322, 322, False),
    ("optical-flow self-attn (N=2048,16 heads,dh=32)", 1, 2048, 2048, 16, 32, 32, False),
    ("optical-flow dec x-attn (O=182528,N=2048,dh=512)", 1, 182528, 2048, 1, 512, 512, False),
    ("perceiver-ar prefix x-attn (N=1024,M=16384,dh=128,causal)", 8, 1024, 16384, 8, 128, 128, True),
    ("perceiver-ar latent self-attn (N=1024,dh=128,causal)", 8, 1024, 1024, 8, 128, 128, True),
    ("perceiver-ar decode step (N=1,M=16384,dh=128)", 8, 1, 16384, 8, 128, 128, True),
    ("sweep M=4096 (N=512,dh=128)", 8, 512, 4096, 8, 128, 128, False),
    ("sweep M=16384", 8, 512, 16384, 8, 128, 128, False),
    ("sweep M=65536 (north star)", 8, 512, 65536, 8, 128, 128, False),
    ("sweep M=262144", 8, 512, 262144, 8, 128, 128, False),
]

rows = []
for name, B, N, M, H, dqk, dv, causal in CASES:
    if ONLY and ONLY not in name:
        continue
    torch.manual_seed(0)
    q = torch.randn(B, N, H * dqk, device="cuda").bfloat16()
    k = torch.randn(B, M, H * dqk, device="cuda").bfloat16()
    v = torch.randn(B, M, H * dv, device="cuda").bfloat16()
    fam = "tcgen05" if ops.tcgen05_supported(q, k, v, H, causal=causal) else "simt"
    if N <= 4 and M >= 1024 and dqk % 8 == 0 and dv % 8 == 0 and max(dqk, dv) <= 256:
        fam = "decode (streaming)"
    elif fam == "tcgen05" and (dqk > 128 or dv > 256):
        fam = "tcgen05 big-head"
    reps = 3 if fam == "simt" and N * M > 1e8 else 20
    for _ in range(2):
        ops.attention(q, k, v, H, dqk ** -0.5, causal=causal)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.attention(q, k, v, H, dqk ** -0.5, causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * H * N * M * (dqk + dv)
    byts = 2.0 * (B * M * H * (dqk + dv) + B * N * H * (dqk + dv))
    rows.append(dict(config=name, kernel=fam, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1)))
    print(json.dumps(rows[-1]), flush=True)
    del q, k, v
    torch.cuda.empty_cache()
print("\n| config | kernel | ms | TFLOP/s (dense count) | algorithmic GB/s |\n|---|---|---|---|---|")
for r in rows:
    print(f"| {r['config']} | {r['kernel']} | {r['ms']} | {r['tflops']} | {r['gbs']} |")
