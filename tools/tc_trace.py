"""Print the in-kernel timeline of CTA 0 (PCV_TRACE=1): per key tile, clock64 deltas for the two softmax
warpgroups and the MMA issuer.  Run on the GPU box: PCV_TRACE=1 PCV_TURNS=0 python tools/tc_trace.py"""
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
print("tile | WG0: S_ready ld max turn exp arrive | WG1: ... | MMA: loop p0_ok pv0 qk0 p1_ok end   (cycles since start)")
for t in range(8, 24):
    row = []
    for r in range(R):
        row.append(" ".join(f"{(x - t0) if x else 0:7d}" for x in a[r][t][:6]))
    print(f"{t:3d} | " + " | ".join(row))
for r, name in enumerate(("WG0", "WG1", "MMA")):
    per = [(a[r][t + 1][0] - a[r][t][0]) for t in range(10, 40)]
    print(name, "period avg", sum(per) / len(per), "min", min(per), "max", max(per))
for r, name in enumerate(("WG0", "WG1")):
    seg = [[a[r][t][e + 1] - a[r][t][e] for e in range(5)] for t in range(10, 40)]
    print(name, "avg phase durations [ld, max, turnwait, exp, st+arrive]:", [sum(s[i] for s in seg) / len(seg) for i in range(5)],
          "wait for next S:", sum(a[r][t + 1][0] - a[r][t][5] for t in range(10, 40)) / 30)
seg = [[a[2][t][e + 1] - a[2][t][e] for e in range(5)] for t in range(10, 40)]
m=a[2]
print("MMA fine: [p1_ok->pv1 issued, commit kv_empty(V), qk1 issue, commit s_full1 + kv_empty(K)]:", [sum(x)/30 for x in zip(*[[m[t][6]-m[t][4], 0, m[t][7]-m[t][6], m[t][5]-m[t][7]] for t in range(10,40)])])
print("MMA avg [wait p0, issue pv0, (wait K)+issue qk0, wait p1, issue pv1+qk1]:", [sum(s[i] for s in seg) / len(seg) for i in range(5)])
