"""Helpers shared by the -m gpu parity tests: the oracle is evaluated in fp64 on the SAME bf16-rounded
operands the kernel sees; tolerance is stated relative to the largest reference magnitude."""
import torch

from oracle import mha_oracle as O


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float64)


def oracle_core(q, k, v, H, scale, pad=None, causal=False):
    """q (Bq,N,H*d), k (B,M,H*d), v (B,M,H*dv) bf16/any on any device -> (B,N,H*dv) fp64 CPU."""
    qc, kc, vc = (t.detach().cpu().to(torch.float64) for t in (q, k, v))
    B = kc.shape[0]
    qh = O.split_heads(qc.expand(B, -1, -1), H)
    out = O.core_attention(qh, O.split_heads(kc, H), O.split_heads(vc, H), scale,
                           None if pad is None else pad.cpu(), causal)
    return O.merge_heads(out)


def assert_close(got, ref, rel, what=""):
    """max |got-ref| <= rel * max|ref|   (rel: 1e-2 for bf16-operand tensor-core paths: P and the output
    are rounded to bf16 = 2^-9 relative each; 6e-3 for the fp32-math SIMT path: output rounding only)."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite values in kernel output"
    err = (got - ref).abs().max().item()
    bound = rel * max(ref.abs().max().item(), 1e-6)
    assert err <= bound, f"{what}: max err {err:.3e} > {bound:.3e}"
    return err


# --------------------------------------------------------------------------------------------------
# The stated parity gate:
#     max|kernel - ref_fp64|  <=  2 * max|ref_bf16_eager - ref_fp64|  +  1e-3 * max|ref_fp64|
# ref_fp64 is the reference algorithm in float64 on the SAME bf16-rounded operands; ref_bf16_eager is
# what the reference's own eager code gives when run in bf16 (q*scale, einsum, masked_fill, softmax, einsum, each
# rounding to bf16 — modules.py:123-167).  Both are evaluated with plain torch ops on the device, so the gate
# is DERIVED per case instead of being a hard-coded relative tolerance.
# --------------------------------------------------------------------------------------------------
def torch_core(q, k, v, H, scale, pad=None, causal=False, dtype=torch.float64):
    """Reference algorithm (modules.py:123-167) with torch ops on q's device in `dtype`.
    q (Bq,N,H*d), k (B,M,H*d), v (B,M,H*dv) -> (B,N,H*dv) in `dtype`."""
    B, M = k.shape[0], k.shape[1]
    N = q.shape[1]
    qh = q.to(dtype).expand(B, -1, -1).reshape(B, N, H, -1).transpose(1, 2)
    kh = k.to(dtype).reshape(B, M, H, -1).transpose(1, 2)
    vh = v.to(dtype).reshape(B, M, H, -1).transpose(1, 2)
    qh = qh * scale                                                      # :124
    attn = torch.einsum("bhic,bhjc->bhij", qh, kh)                       # :151
    neg = -torch.finfo(attn.dtype).max                                   # :152
    if pad is not None:
        attn.masked_fill_(pad.to(q.device).bool()[:, None, None, :], neg)    # :154-155
    if causal:
        cm = torch.ones(N, M, device=q.device, dtype=torch.bool).triu(M - N + 1)  # :135-140
        attn.masked_fill_(cm, neg)                                       # :157-158
    attn = attn.softmax(dim=-1)                                          # :160
    o = torch.einsum("bhij,bhjc->bhic", attn, vh)                        # :163
    return o.transpose(1, 2).reshape(B, N, -1)                           # :166-167


def derived_bound(ref64, eager):
    """(bound, eager_err, ref_max) of the stated gate for one case."""
    ref64 = ref64.double()
    eager_err = (eager.double().to(ref64.device) - ref64).abs().max().item()
    ref_max = ref64.abs().max().item()
    return 2.0 * eager_err + 1e-3 * max(ref_max, 1e-30), eager_err, ref_max


def assert_parity(got, q, k, v, H, scale, pad=None, causal=False, what="", eager_dtype=None, floor=0.0):
    """Check `got` against the fp64 reference with the DERIVED gate; returns (err, bound, eager_err).

    `floor`: lower limit of the bound for paths that add roundings the eager reference does not have (stated by
    the caller where used)."""
    eager_dtype = q.dtype if eager_dtype is None else eager_dtype
    ref = torch_core(q, k, v, H, scale, pad, causal, torch.float64)
    eager = torch_core(q, k, v, H, scale, pad, causal, eager_dtype)
    bound, eager_err, ref_max = derived_bound(ref, eager)
    bound = max(bound, floor * ref_max)
    g = got.detach().double().to(ref.device)
    assert g.shape == ref.shape, (g.shape, ref.shape)
    assert torch.isfinite(g).all(), f"{what}: non-finite values in kernel output"
    err = (g - ref).abs().max().item()
    print(f"[parity] {what}: err {err:.3e}  bound {bound:.3e} (= 2 x eager {eager_err:.3e} + 1e-3 x max|ref| {ref_max:.3e})")
    assert err <= bound, f"{what}: max err {err:.3e} > derived bound {bound:.3e} (eager bf16 err {eager_err:.3e}, max|ref| {ref_max:.3e})"
    return err, bound, eager_err


def torch_cross_attention(sd, x_q, x_kv, H, pad=None, dtype=torch.float64, device="cuda"):
    """CrossAttention.forward (reference modules.py:204-230 -> :113-170) restated with torch ops in `dtype` on
    `device` from a reference state_dict: the fp64 yardstick / the eager-bf16 yardstick of module-level cases."""
    import torch.nn.functional as F

    w = {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}
    xq = x_q.to(device=device, dtype=dtype)
    xkv = x_kv.to(device=device, dtype=dtype)
    xq = F.layer_norm(xq, xq.shape[-1:], w["q_norm.weight"], w["q_norm.bias"], 1e-5)          # :220
    xkv = F.layer_norm(xkv, xkv.shape[-1:], w["kv_norm.weight"], w["kv_norm.bias"], 1e-5)     # :226
    a = "attention."
    q = F.linear(xq, w[a + "q_proj.weight"], w.get(a + "q_proj.bias"))                          # :113
    k = F.linear(xkv, w[a + "k_proj.weight"], w.get(a + "k_proj.bias"))                         # :114
    v = F.linear(xkv, w[a + "v_proj.weight"], w.get(a + "v_proj.bias"))                         # :115
    scale = (q.shape[-1] // H) ** -0.5                                                          # :73
    o = torch_core(q, k, v, H, scale, None if pad is None else pad.to(device), False, dtype)
    return F.linear(o, w[a + "o_proj.weight"], w.get(a + "o_proj.bias"))                        # :168
