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
