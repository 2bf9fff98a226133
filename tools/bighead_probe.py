"""Probe: big-head kernel time per key tile for several (dqk, dv) — which side (K boxes / V boxes / fixed chain) costs what."""
import sys, torch
sys.path.insert(0, ".")
from perceiver_io_b200 import ops
B, N, M, H = 1, 2048, 73728, 1
def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for dqk, dv in ((192, 64), (192, 192), (192, 384), (320, 64), (320, 192), (320, 320), (328, 328), (384, 384), (256, 256)):
    q = torch.randn(B, N, dqk, device="cuda").bfloat16(); k = torch.randn(B, M, dqk, device="cuda").bfloat16(); v = torch.randn(B, M, dv, device="cuda").bfloat16()
    ms = timed(lambda: ops.attention(q, k, v, H, dqk ** -0.5))
    tiles = (M // 128) * (N // 128) / 144   # key tiles per CTA (9 groups x 16 CTAs)
    print(f"dqk {dqk} dv {dv}: {ms:.3f} ms, {ms * 1e3 / tiles:.2f} us per key tile per CTA, {2.0 * B * H * N * M * (dqk + dv) / ms / 1e9:.0f} TFLOP/s")
