"""Per-CTA wall time of the attention kernel (library built with `make TRACE=1`, PCV_TRACE=1): start / end skew,
duration spread and the slowest SMs — separates "steady-state period" from prologue, tail and imbalance."""
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
for _ in range(3):
    ops.attention(q, k, v, H, (d // H) ** -0.5)
torch.cuda.synchronize()
NST, NC = 3 * 48 * 8, 1024
buf = (C.c_uint64 * (NST + 8 * NC))()
lib = _lib.lib()
lib.pcv_debug_trace_read.restype = C.c_int
lib.pcv_debug_trace_read.argtypes = [C.POINTER(C.c_uint64), C.c_int32]
assert lib.pcv_debug_trace_read(buf, NST + 8 * NC) == 0
recs = [tuple(buf[NST + 8 * i + j] for j in range(6)) for i in range(NC)]
recs = [(i, *r) for i, r in enumerate(recs) if r[0]]
t0 = min(r[1] for r in recs)
t1 = max(r[2] for r in recs)
print(f"{len(recs)} CTAs, kernel span {(t1 - t0) / 1e3:.1f} us")
dur = sorted((r[2] - r[1]) / 1e3 for r in recs)
print("CTA duration us: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" %
      (dur[0], dur[len(dur) // 10], dur[len(dur) // 2], dur[9 * len(dur) // 10], dur[-1]))
st = sorted((r[1] - t0) / 1e3 for r in recs)
en = sorted((t1 - r[2]) / 1e3 for r in recs)
print("start skew us: median %.1f max %.1f | idle before kernel end us: median %.1f max %.1f" %
      (st[len(st) // 2], st[-1], en[len(en) // 2], en[-1]))
per_tile = sorted(((r[2] - r[1]) / max(r[4], 1), r[0], r[3], r[4]) for r in recs)
print("ns per key tile: min %.1f median %.1f max %.1f" % (per_tile[0][0], per_tile[len(per_tile) // 2][0], per_tile[-1][0]))
print("slowest CTAs (ns/tile, cta, sm, tiles):", [(round(a, 1), b, c, d) for a, b, c, d in per_tile[-8:]])
print("fastest CTAs (ns/tile, cta, sm, tiles):", [(round(a, 1), b, c, d) for a, b, c, d in per_tile[:8]])
# duration by SM parity / GPC guess: print a coarse histogram
import collections
h = collections.Counter(int(x[0] // 20) * 20 for x in per_tile)
print("histogram ns/tile (bucket 20 ns):", sorted(h.items()))
mhz = sorted((r[6] - r[5]) / max(r[2] - r[1], 1) * 1e3 for r in recs)
print("SM clock during the kernel (clock64 cycles / globaltimer ns), MHz: min %.0f median %.0f max %.0f" % (mhz[0], mhz[len(mhz) // 2], mhz[-1]))
cyc = sorted((r[6] - r[5]) / max(r[4], 1) for r in recs)
print("cycles per key tile (whole CTA): min %.0f median %.0f max %.0f" % (cyc[0], cyc[len(cyc) // 2], cyc[-1]))
