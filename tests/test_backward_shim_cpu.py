"""CPU check of the training-support backward shim (ops._FusedAttention.backward): the chunked flash-attention
recurrence, fed with saved operands only, must reproduce torch autograd of the reference's eager formula
(modules.py:146-164: masked_fill_ blocks gradients through filled scores, a fully padded row is uniform) while never
holding more than `backward_config["max_score_bytes"]` of scores."""
import pytest
import torch

from perceiver_io_b200 import ops


def _eager(q, k, v, H, scale, pad, causal):
    B, M, N = k.shape[0], k.shape[1], q.shape[1]
    qh = q.expand(B, -1, -1).reshape(B, N, H, -1).transpose(1, 2) * scale
    kh = k.reshape(B, M, H, -1).transpose(1, 2)
    vh = v.reshape(B, M, H, -1).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    neg = -torch.finfo(s.dtype).max
    if pad is not None:
        s = s.masked_fill(pad[:, None, None, :], neg)
    if causal:
        s = s.masked_fill(torch.ones(N, M, dtype=torch.bool).triu(M - N + 1), neg)
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, N, -1)


class _Ctx:
    pass


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("with_stats", [False, True])
def test_chunked_backward_matches_autograd(causal, with_stats, monkeypatch):
    B, N, M, H, d = 2, 5, 300, 2, 8
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, N, H * d, generator=g, dtype=torch.float64, requires_grad=True)
    k = torch.randn(B, M, H * d, generator=g, dtype=torch.float64, requires_grad=True)
    v = torch.randn(B, M, H * d, generator=g, dtype=torch.float64, requires_grad=True)
    pad = torch.zeros(B, M, dtype=torch.bool)
    pad[0, :50] = True
    pad[1, :] = True
    scale = d ** -0.5
    o = _eager(q, k, v, H, scale, pad, causal)
    go = torch.randn(o.shape, generator=g, dtype=torch.float64)
    gq, gk, gv = torch.autograd.grad(o, (q, k, v), go)

    monkeypatch.setattr(ops, "_compute_dtype", lambda dt: torch.float32)
    monkeypatch.setitem(ops.backward_config, "max_score_bytes", 4 * B * H * N * 128)   # forces 3 key chunks
    pm = pl = None
    if with_stats:  # the statistics the forward kernel would have saved (log2 domain)
        qh = q.detach().float().expand(B, -1, -1).reshape(B, N, H, -1).transpose(1, 2)
        kh = k.detach().float().reshape(B, M, H, -1).transpose(1, 2)
        t = (qh @ kh.transpose(-1, -2)) * (scale * 1.4426950408889634)
        neg = -torch.finfo(torch.float32).max
        t = t.masked_fill(pad[:, None, None, :], neg)
        if causal:
            t = t.masked_fill(torch.ones(N, M, dtype=torch.bool).triu(M - N + 1), neg)
        pm = t.amax(-1) - 3.0                       # any reference maximum works, not only the true one
        pl = torch.exp2(t - pm[..., None]).sum(-1)
    ctx = _Ctx()
    ctx.saved_tensors = (q.detach().float(), k.detach().float(), v.detach().float(), pad, o.detach().float(), pm, pl)
    ctx.meta = (H, scale, causal)
    r = ops._FusedAttention.backward(ctx, go.float())
    for got, ref, name in zip(r[:3], (gq, gk, gv), "qkv"):
        assert got.shape == ref.shape
        assert (got.double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), name
    assert all(x is None for x in r[3:])


def test_backward_dispatch_rules(monkeypatch):
    """Which backward runs: an unknown mode is an error; the shim cannot regenerate a dropout mask, so a dropout forward
    whose gradients cannot go through the kernels must fail loudly instead of returning dropout-free gradients."""
    B, N, M, H, d = 1, 4, 64, 1, 8
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, n, H * d, generator=g) for n in (N, M, M))
    o = _eager(q.double(), k.double(), v.double(), H, d ** -0.5, None, False).float()
    ctx = _Ctx()
    ctx.saved_tensors = (q, k, v, None, o, None, None)
    ctx.meta = (H, d ** -0.5, False)
    monkeypatch.setattr(ops, "_compute_dtype", lambda dt: torch.float32)
    monkeypatch.setitem(ops.backward_config, "impl", "fastest")
    with pytest.raises(ValueError, match="backward_config"):
        ops._FusedAttention.backward(ctx, torch.ones_like(o))
    monkeypatch.setitem(ops.backward_config, "impl", "auto")
    r = ops._FusedAttention.backward(ctx, torch.ones_like(o))     # CPU tensors, no statistics: the shim
    assert r[0].shape == q.shape and len(r) == 10 and all(x is None for x in r[3:])
    ctx.dropout = (0.1, 7)
    with pytest.raises(RuntimeError, match="dropout"):
        ops._FusedAttention.backward(ctx, torch.ones_like(o))
