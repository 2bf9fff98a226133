"""Minimal launch sequence for ncu: optical-flow encoder cross-attention geometry (dh = 322), reduced M."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import ops
B, N, M, H, d = 1, 2048, int(os.environ.get("PCV_M", 36864)), 1, int(os.environ.get("PCV_D", 322))
torch.manual_seed(0)
q = torch.randn(B, N, H * d, device="cuda").bfloat16()
k = torch.randn(B, M, H * d, device="cuda").bfloat16()
v = torch.randn(B, M, H * d, device="cuda").bfloat16()
for _ in range(3):
    out = ops.attention(q, k, v, H, d ** -0.5)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
