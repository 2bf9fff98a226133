"""Latent self-attention stack of the MLM config (BASELINE.json configs[1]; reference scripts/text/mlm.py:16-39):
26 layers, N = 256 latents, D = 1280, 8 heads (32 / 160 channels per head), bf16, batch 8.  Per-forward latency of
(a) library projections (LayerNorm + 3 x nn.Linear per layer), (b) the LayerNorm-folded one-GEMM QKV projection +
tcgen05 o_proj, (c) the same recorded as ONE CUDA graph (perceiver_io_b200.graphs)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import perceiver_io_b200 as P  # noqa: E402
from perceiver_io_b200 import _lib, modules  # noqa: E402
from perceiver_io_b200.graphs import graph_latent_block  # noqa: E402

B, N, D, H, L = 8, 256, 1280, 8, 26
torch.manual_seed(0)
block = P.SelfAttentionBlock(num_layers=L, num_heads=H, num_channels=D, num_qk_channels=256, num_v_channels=1280,
                             widening_factor=1).cuda().bfloat16().eval()
x = torch.randn(B, N, D, device="cuda").bfloat16()


def timed(fn, iters=20, warm=3):
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


res = {"shape": {"B": B, "N": N, "D": D, "H": H, "layers": L}}
with torch.no_grad():
    modules.kv_producer_config["enabled"] = False
    res["library_projections_ms"] = round(timed(lambda: block(x)), 4)
    modules.kv_producer_config["enabled"] = True
    modules.kv_producer_config["min_rows_latent"] = 512
    l0 = _lib.launch_count()
    res["fused_projections_ms"] = round(timed(lambda: block(x)), 4)
    res["library_kernels_per_forward"] = (_lib.launch_count() - l0) // 23
modules.kv_producer_config["min_rows_latent"] = 4096
fast = graph_latent_block(block, x)
res["cuda_graph_ms"] = round(timed(lambda: fast(x)), 4)
print(json.dumps(res))
