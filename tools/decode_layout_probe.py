"""Probe: decode-core bandwidth with token-major (B, M, H*d) vs head-major (B, H, M, d) caches (same kernel)."""
import sys, torch
sys.path.insert(0, ".")
from perceiver_io_b200 import ops
B, L, C, H = 8, 16384, 1024, 8
d = C // H
q = torch.randn(B, 1, C, device="cuda").bfloat16()
def timed(fn, iters=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
byts = 2.0 * B * L * C * 2
for layout in ("token_major", "head_major"):
    caches = []
    for _ in range(4):
        if layout == "token_major":
            caches.append((torch.randn(B, L, C, device="cuda").bfloat16(), torch.randn(B, L, C, device="cuda").bfloat16()))
        else:
            caches.append((torch.randn(B, H, L, d, device="cuda").bfloat16().permute(0, 2, 1, 3), torch.randn(B, H, L, d, device="cuda").bfloat16().permute(0, 2, 1, 3)))
    for impl in ("decode", "tcgen05"):
        it = [0]
        def step():
            k, v = caches[it[0] % 4]; it[0] += 1
            return ops.attention(q, k, v, H, d ** -0.5, causal=True, impl=impl)
        ms = timed(step)
        from perceiver_io_b200 import _lib
        torch.cuda.synchronize(); _lib.profile_begin()
        for _ in range(20): step()
        torch.cuda.synchronize(); tot, n = _lib.profile_end()
        print(layout, impl, round(ms, 4), "ms/step wall", round(byts / ms / 1e6), "GB/s | kernel alone", round(tot / n, 4), "ms", round(byts / (tot / n) / 1e6), "GB/s")
    del caches
# plain read bandwidth of the same bytes (sum reduction) for reference
x = torch.randn(B, L, C, device="cuda").bfloat16(); y = torch.randn(B, L, C, device="cuda").bfloat16()
ms = timed(lambda: (x.sum(), y.sum()))
print("torch sum over the same bytes", round(ms, 4), "ms", round(byts / ms / 1e6), "GB/s")
