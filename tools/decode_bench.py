"""Perceiver-AR decode shape (B=8, 16 k cached tokens, C=1024, H=8; BASELINE.json configs[3]):
  (1) the attention core for ONE new token against the cache — streaming decode kernel vs the tcgen05 kernel —
      as achieved HBM GB/s (algorithmic bytes = B*M*(Dqk+Dv)*2) against the measured copy bandwidth;
  (2) per-token latency of one cached attention layer with arena-backed caches (in-place append) vs concat."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import perceiver_io_b200 as P  # noqa: E402
from perceiver_io_b200 import ops  # noqa: E402

B, L, C, H, STEPS = 8, 16384, 1024, 8, 64
torch.manual_seed(0)
peak = 6587.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timed(fn, iters=50, warm=5):
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


res = {"shape": {"B": B, "cached_tokens": L, "C": C, "H": H}, "hbm_peak_gbs": peak}
# (1) core: several independent caches so that consecutive launches do not find K/V in the 126 MB L2
caches = [(torch.randn(B, L, C, device="cuda").bfloat16(), torch.randn(B, L, C, device="cuda").bfloat16()) for _ in range(4)]
q = torch.randn(B, 1, C, device="cuda").bfloat16()
byts = 2.0 * B * L * C * 2
for impl in ("decode", "tcgen05"):
    it = [0]

    def step():
        k, v = caches[it[0] % 4]
        it[0] += 1
        return ops.attention(q, k, v, H, (C // H) ** -0.5, causal=True, impl=impl)

    ms = timed(step)
    res[f"core_{impl}_ms"] = round(ms, 4)
    res[f"core_{impl}_gbs"] = round(byts / ms / 1e6, 1)
    res[f"core_{impl}_frac_of_hbm_peak"] = round(byts / ms / 1e6 / peak, 3)
del caches

# (2) one cached attention layer, one token per step
mha = P.MultiHeadAttention(num_heads=H, num_q_input_channels=C, num_kv_input_channels=C, causal_attention=True)
mha = mha.cuda().bfloat16().eval()
prompt = torch.randn(B, L, C, device="cuda").bfloat16()
tok = torch.randn(B, 1, C, device="cuda").bfloat16()
for enabled in (True, False):
    ops.kv_arena_config["enabled"] = enabled
    with torch.no_grad():
        empty = (torch.zeros(B, 0, C, device="cuda", dtype=torch.bfloat16),) * 2
        cache = mha(prompt[:, -1:], prompt, kv_cache=empty).kv_cache
        for _ in range(4):
            cache = mha(tok, tok, kv_cache=cache).kv_cache
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(STEPS):
            out = mha(tok, tok, kv_cache=cache)
            cache = out.kv_cache
        e1.record()
        torch.cuda.synchronize()
    res["layer_ms_per_token_arena" if enabled else "layer_ms_per_token_concat"] = round(e0.elapsed_time(e1) / STEPS, 4)
    del cache
    torch.cuda.empty_cache()
print(json.dumps(res))
