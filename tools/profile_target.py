"""Minimal launch sequence for ncu: the north-star attention core, a few launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import ops  # noqa: E402

B, N, M, d, H = 8, 512, int(os.environ.get("PCV_M", 65536)), 1024, 8
torch.manual_seed(0)
q = torch.randn(B, N, d, device="cuda").bfloat16()
k = torch.randn(B, M, d, device="cuda").bfloat16()
v = torch.randn(B, M, d, device="cuda").bfloat16()
for _ in range(int(os.environ.get("PCV_ITERS", 3))):
    out = ops.attention(q, k, v, H, (d // H) ** -0.5)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
