"""Stress the backward / dropout kernels on small shapes (few CTAs, one or two tiles each: the hand-off corner cases) and
report the watchdog record if a launch fails.  python tools/bwd_stress.py [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceiver_io_b200 import _lib, ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
CASES = [(2, 96, 352, 2, 64, 64, True), (1, 128, 128, 1, 64, 64, False), (1, 100, 40, 2, 32, 96, False),
         (2, 130, 700, 2, 128, 64, False), (3, 300, 900, 2, 64, 128, True), (2, 256, 1024, 2, 128, 128, False)]
g = torch.Generator(device="cuda").manual_seed(0)
try:
    for case in CASES:
        B, N, M, H, dqk, dv, causal = case
        q = torch.randn(B, N, H * dqk, device="cuda", generator=g).bfloat16()
        k = torch.randn(B, M, H * dqk, device="cuda", generator=g).bfloat16()
        v = torch.randn(B, M, H * dv, device="cuda", generator=g).bfloat16()
        go = torch.randn(B, N, H * dv, device="cuda", generator=g).bfloat16()
        scale = dqk ** -0.5
        ref = None
        for it in range(iters):
            po, pm, pl = ops.attention_partial(q, k, v, H, scale, causal=causal)
            out = ops.combine_partials(po[None], pm[None], pl[None], q.dtype)
            drop = 0.1 if it % 2 else 0.0
            if drop:
                out = ops.attention_dropout_forward(q, k, v, pm, pl, H, scale, drop, 99, causal=causal)
            grads = ops.attention_backward(q, k, v, out, go, pm, pl, H, scale, causal=causal, dropout_p=drop, dropout_seed=99)
            if it < 2:
                torch.cuda.synchronize()
                ref = ref or {}
                ref[it] = [t.clone() for t in grads]
            elif it % 50 < 2:
                torch.cuda.synchronize()
                for a, b_ in zip(grads, ref[it % 2]):
                    assert torch.equal(a[..., :8], b_[..., :8]) or (a.float() - b_.float()).abs().max().item() < 1e-2
        torch.cuda.synchronize()
        print("ok", case, flush=True)
except Exception as e:  # noqa: BLE001
    print("FAILED", case, "iteration", it, type(e).__name__, str(e)[:200])
    print("watchdog record:", _lib.debug_read() if hasattr(_lib, "debug_read") else None)
    sys.exit(1)
print("stress ok")
