"""Minimal launch sequence for ncu: the fused K/V producer at a reduced row count (PCV_ROWS, default 131072)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import ops  # noqa: E402

rows, C, n = int(os.environ.get("PCV_ROWS", 131072)), 1024, 1024
torch.manual_seed(0)
x = torch.randn(rows, C, device="cuda").bfloat16()
w = [torch.randn(n, C, device="cuda").bfloat16() * C ** -0.5 for _ in range(2)]
w_cat, col_st = ops.fold_ln_linear(torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), w, [None, None], torch.bfloat16)
for cg in (2, 1, 2, 1):
    k, v = ops.kv_project(x, w_cat, col_st, n, n, eps=1e-5, cta_group=cg)
torch.cuda.synchronize()
print("ok", float(k.float().abs().mean()))
