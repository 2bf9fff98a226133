"""Per-token latency of one cached attention layer at the Perceiver-AR decode shape (B=8, 16 k cached tokens,
C=1024, H=8): ops.kv_append with arena-backed caches (in-place append) vs plain concat every step."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import perceiver_io_b200 as P  # noqa: E402
from perceiver_io_b200 import ops  # noqa: E402

B, L, C, H, STEPS = 8, 16384, 1024, 8, 64
torch.manual_seed(0)
mha = P.MultiHeadAttention(num_heads=H, num_q_input_channels=C, num_kv_input_channels=C, causal_attention=True)
mha = mha.cuda().bfloat16().eval()
prompt = torch.randn(B, L, C, device="cuda").bfloat16()
tok = torch.randn(B, 1, C, device="cuda").bfloat16()
res = {}
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
    res["arena" if enabled else "concat"] = e0.elapsed_time(e1) / STEPS
    del cache
    torch.cuda.empty_cache()
print(json.dumps({"shape": {"B": B, "cached_tokens": L, "C": C, "H": H}, "ms_per_token_arena": round(res["arena"], 4),
                  "ms_per_token_concat": round(res["concat"], 4), "speedup": round(res["concat"] / res["arena"], 2)}))
