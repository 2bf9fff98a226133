"""-m gpu parity tests of the fused K/V producer (pcv_ln_stats + pcv_kv_project) against the
reference arithmetic  k_proj(kv_norm(x)), v_proj(kv_norm(x))  (reference modules.py:226, :114-115) in float64 on
the same bf16/fp16 operands.  The gate is derived per case like the attention gate (gpu_util):
    max|kernel - ref_fp64| <= 2 * max|ref_eager - ref_fp64| + 1e-3 * max|ref_fp64|
where ref_eager is torch's own LayerNorm -> Linear run eagerly in the operand dtype on the device."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import derived_bound

pytestmark = pytest.mark.gpu


def _case(rows, C, n_k, n_v, dtype, seed=0, mean=0.0, affine=True, bias=True):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(rows, C, generator=g) * 1.3 + mean).to(dtype).cuda()
    gamma = (1.0 + 0.2 * torch.randn(C, generator=g)).cuda() if affine else torch.ones(C).cuda()
    beta = (0.3 * torch.randn(C, generator=g)).cuda() if affine else torch.zeros(C).cuda()
    wk = (torch.randn(n_k, C, generator=g) * C ** -0.5).cuda()
    wv = (torch.randn(n_v, C, generator=g) * C ** -0.5).cuda()
    bk = (0.1 * torch.randn(n_k, generator=g)).cuda() if bias else None
    bv = (0.1 * torch.randn(n_v, generator=g)).cuda() if bias else None
    # the module holds its parameters in the compute dtype (.bfloat16() model): round them once, like the reference
    r = lambda t: None if t is None else t.to(dtype)
    return x, r(gamma), r(beta), r(wk), r(bk), r(wv), r(bv)


def _reference(x, gamma, beta, wk, bk, wv, bv, dtype, eps=1e-5):
    C = x.shape[-1]
    c = lambda t: None if t is None else t.to(dtype)
    xn = F.layer_norm(x.to(dtype), (C,), c(gamma), c(beta), eps)
    return F.linear(xn, c(wk), c(bk)), F.linear(xn, c(wv), c(bv))


def _check(got, ref64, eager, what):
    bound, eager_err, ref_max = derived_bound(ref64, eager)
    err = (got.double() - ref64).abs().max().item()
    print(f"[parity] {what}: err {err:.3e} bound {bound:.3e} (eager {eager_err:.3e}, max|ref| {ref_max:.3e})")
    assert torch.isfinite(got).all(), what
    assert err <= bound, f"{what}: err {err:.3e} > derived bound {bound:.3e} (eager {eager_err:.3e}, max|ref| {ref_max:.3e})"


SHAPES = [
    # rows, C, n_k, n_v
    (1000, 1024, 1024, 1024),   # north-star widths, ragged row count (partial last row tile, both CTA variants)
    (4096, 768, 256, 1280),     # MLM encoder: C=768 (12 k-blocks), 32/160 head widths -> 6 column tiles
    (300, 64, 64, 72),          # one k-block, fewer columns than one tile, V width not a multiple of 64
    (513, 72, 128, 8),          # C not a multiple of 64 (TMA zero fill of the K tail), tiny V
    (2048, 512, 512, 0),        # K only
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("cg", [1, 2], ids=["cta1", "cta2"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
@pytest.mark.parametrize("stats", ["fused", "separate"])
def test_kv_project_matches_layernorm_linear(shape, cg, dtype, stats):
    from perceiver_io_b200 import ops

    rows, C, n_k, n_v = shape
    x, gamma, beta, wk, bk, wv, bv = _case(rows, C, n_k, n_v, dtype, seed=rows, mean=0.7)
    ws = [wk] + ([wv] if n_v else [])
    bs = [bk] + ([bv] if n_v else [])
    w_cat, col_st = ops.fold_ln_linear(gamma, beta, ws, bs, dtype)
    k, v = ops.kv_project(x, w_cat, col_st, n_k, n_v, eps=1e-5, cta_group=cg, stats=stats)
    rk64, rv64 = _reference(x, gamma, beta, wk, bk, wv if n_v else wk[:0], bv if n_v else None, torch.float64)
    ek, ev = _reference(x, gamma, beta, wk, bk, wv if n_v else wk[:0], bv if n_v else None, dtype)
    assert k.shape == (rows, n_k) and k.dtype == dtype
    _check(k, rk64, ek, f"K {shape} cg={cg}")
    if n_v:
        assert v.shape == (rows, n_v)
        _check(v, rv64, ev, f"V {shape} cg={cg}")
    else:
        assert v is None


def test_ln_stats_matches_torch():
    from perceiver_io_b200 import ops

    g = torch.Generator().manual_seed(1)
    for rows, C in ((257, 1024), (33, 72), (5, 131)):   # 131: scalar path (odd width, unaligned rows)
        x = (torch.randn(rows, C, generator=g) * 2 + 5).bfloat16().cuda()
        st = ops.ln_stats(x, 1e-5).double()
        xd = x.double()
        mean = xd.mean(1)
        rstd = (xd.var(1, unbiased=False) + 1e-5).rsqrt()
        assert (st[:, 0] - mean).abs().max().item() <= 1e-5 * mean.abs().max().item()
        assert ((st[:, 1] - rstd) / rstd).abs().max().item() <= 1e-5


def test_large_row_mean_does_not_break_the_folded_cancellation():
    """|mean| >> std: rstd * (x.w' - mean * s) cancels almost completely; s is the row sum of the ROUNDED weights,
    so what is left is fp32 accumulation error amplified by mean/std (here 200) — far below the bf16 output step."""
    from perceiver_io_b200 import ops

    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(3)
    rows, C, n = 640, 1024, 256
    # values on a coarse grid so that x is exactly representable in bf16 although its mean is large
    x = (torch.randint(-8, 9, (rows, C), generator=g).float() * 0.25 + 48.0).to(dtype).cuda()
    _, gamma, beta, wk, bk, wv, bv = _case(rows, C, n, n, dtype, seed=4)
    w_cat, col_st = ops.fold_ln_linear(gamma, beta, [wk, wv], [bk, bv], dtype)
    rk64, rv64 = _reference(x, gamma, beta, wk, bk, wv, bv, torch.float64)
    ek, ev = _reference(x, gamma, beta, wk, bk, wv, bv, dtype)
    for stats in ("fused", "separate"):   # one-pass shifted statistics in the GEMM kernel / two-pass pcv_ln_stats
        k, v = ops.kv_project(x, w_cat, col_st, n, n, stats=stats)
        _check(k, rk64, ek, f"K large mean ({stats})")
        _check(v, rv64, ev, f"V large mean ({stats})")


def test_plain_projection_without_layernorm():
    from perceiver_io_b200 import ops

    dtype = torch.bfloat16
    x, _, _, wk, bk, wv, bv = _case(700, 256, 192, 64, dtype, seed=9)
    w_cat, col_st = ops.fold_ln_linear(None, None, [wk, wv], [bk, bv], dtype)
    k, v = ops.kv_project(x, w_cat, col_st, 192, 64, eps=None)
    _check(k, F.linear(x.double(), wk.double(), bk.double()), F.linear(x, wk, bk), "K plain")
    _check(v, F.linear(x.double(), wv.double(), bv.double()), F.linear(x, wv, bv), "V plain")


def test_cross_attention_module_fused_producer_equals_library_path():
    """CrossAttention.forward (eval, bf16): the fused producer path against the LayerNorm + nn.Linear path of the
    same module, and both against the fp64 evaluation of the reference formula; checks the cache of folded weights
    is rebuilt after an in-place parameter update."""
    import perceiver_io_b200 as P
    from perceiver_io_b200 import modules

    torch.manual_seed(0)
    B, N, M, D, H = 2, 192, 1536, 512, 4
    layer = P.CrossAttention(num_heads=H, num_q_input_channels=D, num_kv_input_channels=D).cuda().bfloat16().eval()
    with torch.no_grad():
        layer.kv_norm.weight.add_(0.1 * torch.randn_like(layer.kv_norm.weight))
        layer.kv_norm.bias.add_(0.1 * torch.randn_like(layer.kv_norm.bias))
    x_q = torch.randn(1, N, D, device="cuda").bfloat16()
    x_kv = (torch.randn(B, M, D, device="cuda") + 0.5).bfloat16()
    pad = torch.zeros(B, M, dtype=torch.bool, device="cuda")
    pad[1, 1000:] = True

    def run(enabled):
        modules.kv_producer_config["enabled"] = enabled
        try:
            with torch.no_grad():
                empty = (torch.empty(B, 0, D, device="cuda", dtype=torch.bfloat16),
                         torch.empty(B, 0, D, device="cuda", dtype=torch.bfloat16))
                out = layer(x_q, x_kv, pad_mask=pad, kv_cache=empty)
            return out
        finally:
            modules.kv_producer_config["enabled"] = True

    fused, plain = run(True), run(False)
    assert "_pcv_kv_fold" in layer.__dict__
    scale = fused.last_hidden_state.float().abs().max().item()
    assert (fused.last_hidden_state.float() - plain.last_hidden_state.float()).abs().max().item() <= 2e-2 * scale
    # the cache entries are the un-rotated K / V rows (reference modules.py:117-121): compare them directly
    xn = F.layer_norm(x_kv.double(), (D,), layer.kv_norm.weight.double(), layer.kv_norm.bias.double(), layer.kv_norm.eps)
    k64 = F.linear(xn, layer.attention.k_proj.weight.double(), layer.attention.k_proj.bias.double())
    _check(fused.kv_cache[0], k64, plain.kv_cache[0], "module K rows")
    # in-place update of a folded parameter invalidates the cache
    key0 = layer.__dict__["_pcv_kv_fold"][0]
    with torch.no_grad():
        layer.attention.k_proj.weight.mul_(1.5)
    fused2 = run(True)
    assert layer.__dict__["_pcv_kv_fold"][0] != key0
    k64b = F.linear(xn, layer.attention.k_proj.weight.double(), layer.attention.k_proj.bias.double())
    _check(fused2.kv_cache[0], k64b, run(False).kv_cache[0], "module K rows after update")


def test_training_mode_uses_the_autograd_capable_path():
    import perceiver_io_b200 as P

    torch.manual_seed(1)
    layer = P.CrossAttention(num_heads=2, num_q_input_channels=128, num_kv_input_channels=128).cuda().bfloat16().train()
    x_q = torch.randn(1, 16, 128, device="cuda").bfloat16()
    x_kv = torch.randn(2, 1024, 128, device="cuda").bfloat16()
    out = layer(x_q, x_kv).last_hidden_state
    out.float().square().mean().backward()
    assert layer.attention.k_proj.weight.grad is not None and "_pcv_kv_fold" not in layer.__dict__
