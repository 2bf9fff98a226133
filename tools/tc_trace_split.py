"""In-kernel timeline of CTA 0 for the split-tile kernel (library built with `make TRACE=1`, PCV_TRACE=1):
per key tile, clock64 stamps of the two softmax warpgroups (half A / half B) and of the issuer of query tile 0."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCV_TRACE", "1")
from perceiver_io_b200 import _lib, ops  # noqa: E402

B, N, M, d, H = 8, 512, 65536, 1024, 8
torch.manual_seed(0)
q = torch.randn(B, N, d, device="cuda").bfloat16()
k = torch.randn(B, M, d, device="cuda").bfloat16()
v = torch.randn(B, M, d, device="cuda").bfloat16()
for _ in range(2):
    ops.attention(q, k, v, H, (d // H) ** -0.5)
torch.cuda.synchronize()
R, T, E = 3, 48, 8
buf = (C.c_uint64 * (R * T * E))()
lib = _lib.lib()
lib.pcv_debug_trace_read.restype = C.c_int
lib.pcv_debug_trace_read.argtypes = [C.POINTER(C.c_uint64), C.c_int32]
assert lib.pcv_debug_trace_read(buf, R * T * E) == 0
a = [[[buf[(r * T + t) * E + e] for e in range(E)] for t in range(T)] for r in range(R)]
t0 = min(x for r in a for t in r for x in t if x)
WG = (0, 1, 4, 5)
print("tile | WG0: A_ready A_done B_ready B_done | WG1: ... | issuer0: top pA_ok A_issued pB_ok end   (cycles since start)")
for t in range(8, 20):
    row = []
    for r in range(2):
        row.append(" ".join(f"{(a[r][t][e] - t0) if a[r][t][e] else 0:7d}" for e in WG))
    row.append(" ".join(f"{(x - t0) if x else 0:7d}" for x in a[2][t][:5]))
    print(f"{t:3d} | " + " | ".join(row))
rng = range(10, 40)
for r, name in enumerate(("WG0", "WG1", "issuer0")):
    per = [(a[r][t + 1][0] - a[r][t][0]) for t in rng]
    print(name, "period avg", sum(per) / len(per), "min", min(per), "max", max(per))
for r, name in enumerate(("WG0", "WG1")):
    x = a[r]
    ph = [[x[t][1] - x[t][0], x[t][4] - x[t][1], x[t][5] - x[t][4], x[t + 1][0] - x[t][5]] for t in rng]
    print(name, "avg [half A softmax, wait S_B, half B softmax, wait next S_A]:", [round(sum(s[i] for s in ph) / len(ph), 1) for i in range(4)])
x = a[2]
ph = [[x[t][1] - x[t][0], x[t][2] - x[t][1], x[t][3] - x[t][2], x[t][4] - x[t][3], x[t + 1][0] - x[t][4]] for t in rng]
print("issuer0 avg [wait P_A, issue PV_A + QK_A(next) + commit, wait P_B, issue PV_B + QK_B(next) + commits, wait V]:",
      [round(sum(s[i] for s in ph) / len(ph), 1) for i in range(5)])
