"""GPU box: fused K/V producer (pcv_ln_stats + pcv_kv_project) vs LayerNorm + two cuBLAS GEMMs at the north-star
module shape (rows = B*M = 524288, C = 1024, n_k = n_v = 1024), CTA-pair and single-CTA variants.
Prints ms, TFLOP/s (4*rows*C*n flops for both projections) and the max deviation from the library path."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from perceiver_io_b200 import ops  # noqa: E402

rows, C, n = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (524288, 1024, 1024)))
dtype = torch.bfloat16
torch.manual_seed(0)
x = (torch.randn(rows, C, device="cuda") + 0.3).to(dtype)
gamma = (1 + 0.1 * torch.randn(C, device="cuda")).to(dtype)
beta = (0.1 * torch.randn(C, device="cuda")).to(dtype)
wk = (torch.randn(n, C, device="cuda") * C ** -0.5).to(dtype)
wv = (torch.randn(n, C, device="cuda") * C ** -0.5).to(dtype)
bk = (0.1 * torch.randn(n, device="cuda")).to(dtype)
bv = (0.1 * torch.randn(n, device="cuda")).to(dtype)
w_cat, col_st = ops.fold_ln_linear(gamma, beta, [wk, wv], [bk, bv], dtype)
flops = 4.0 * rows * C * n


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def lib():
    xn = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    return F.linear(xn, wk, bk), F.linear(xn, wv, bv)


def gemm_only():
    return F.linear(x, wk, bk), F.linear(x, wv, bv)


kl, vl = lib()
for cg, mode in ((2, "fused"), (2, "separate"), (1, "fused"), (1, "separate")):
    k, v = ops.kv_project(x, w_cat, col_st, n, n, eps=1e-5, cta_group=cg, stats=mode)
    torch.cuda.synchronize()
    dk = (k.float() - kl.float()).abs().max().item()
    dv = (v.float() - vl.float()).abs().max().item()
    ms = timed(lambda: ops.kv_project(x, w_cat, col_st, n, n, eps=1e-5, cta_group=cg, stats=mode))
    st = timed(lambda: ops.ln_stats(x, 1e-5))
    print(f"producer cg={cg} stats={mode}: {ms:.3f} ms total ({flops / ms / 1e9:.0f} TFLOP/s); separate ln_stats pass alone {st:.3f} ms "
          f"({rows * C * 2 / st / 1e6:.0f} GB/s); max|dK| {dk:.3e} max|dV| {dv:.3e} vs library (max|K| {kl.float().abs().max().item():.2f})")
ms = timed(lib)
print(f"library LN + 2 x cuBLAS: {ms:.3f} ms ({flops / ms / 1e9:.0f} TFLOP/s)")
ms = timed(gemm_only)
print(f"library 2 x cuBLAS only: {ms:.3f} ms ({flops / ms / 1e9:.0f} TFLOP/s)")
wcat2 = torch.cat([wk, wv])
ms = timed(lambda: F.linear(x, wcat2))
print(f"library 1 x cuBLAS (N=2n): {ms:.3f} ms ({flops / ms / 1e9:.0f} TFLOP/s)")
