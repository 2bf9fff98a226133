"""ncu target: attention backward (pcv_attn_bwd) at the north-star shape, three calls."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import ops  # noqa: E402

B, N, M, H, d = 8, 512, int(os.environ.get("PCV_M", 65536)), 8, 128
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, N, H * d, device="cuda", generator=g).to(torch.bfloat16)
k = torch.randn(B, M, H * d, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(B, M, H * d, device="cuda", generator=g).to(torch.bfloat16)
go = torch.randn(B, N, H * d, device="cuda", generator=g).to(torch.bfloat16)
scale = d ** -0.5
po, pm, pl = ops.attention_partial(q, k, v, H, scale)
out = ops.combine_partials(po[None], pm[None], pl[None], q.dtype)
for _ in range(3):
    ops.attention_backward(q, k, v, out, go, pm, pl, H, scale)
torch.cuda.synchronize()
print("done")
