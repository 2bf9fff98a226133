"""Stage timing of the M-sharded step (torchrun, one rank per GPU): partial kernel / barrier / peer merge / barrier."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceiver_io_b200 import ops  # noqa: E402
from perceiver_io_b200.dist import PeerMerger, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, N, M, d, H = 8, 512, 65536, 1024, 8
m0, m1 = shard_bounds(M, world, rank)
torch.manual_seed(rank)
q = torch.randn(B, N, d, device=dev).bfloat16()
k = torch.randn(B, m1 - m0, d, device=dev).bfloat16()
v = torch.randn(B, m1 - m0, d, device=dev).bfloat16()
pm = PeerMerger.get(B, H, N, d // H, torch.bfloat16, dev, None)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = [0.0] * 5
iters = 30
import ctypes as C
from perceiver_io_b200 import _lib
from perceiver_io_b200.ops import _pcv_dtype, _stream


def merge_kernel_only():
    p = _lib.PeerCombineParams()
    for g in range(pm.world):
        base = pm.part_ptrs[g]
        p.part_o[g] = base
        p.part_m[g] = base + pm.rows * pm.dv * 4
        p.part_l[g] = base + (pm.rows * pm.dv + pm.rows) * 4
        p.out[g] = pm.out_ptrs[g]
    p.o_stride_b, p.o_stride_n, p.o_stride_h = pm.N * pm.H * pm.dv, pm.H * pm.dv, pm.dv
    p.row_begin, p.row_end = pm.row_begin, pm.row_end
    p.num_peers, p.rank = pm.world, pm.rank
    p.B, p.H, p.N, p.dv = pm.B, pm.H, pm.N, pm.dv
    p.dtype = _pcv_dtype(pm.dtype)
    _lib.check(_lib.lib().pcv_attn_combine_peers(C.byref(p), _stream()), "combine_peers")


for it in range(iters + 5):
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    ops.attention_partial(q, k, v, H, (d // H) ** -0.5, m_total=M, m_offset=m0, out=(pm.po, pm.pm, pm.pl))
    ev[1].record()
    pm.part_hdl.barrier(channel=0)
    ev[2].record()
    merge_kernel_only()
    ev[3].record()
    pm.out_hdl.barrier(channel=1)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 5:
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        acc[4] += ev[0].elapsed_time(ev[4])
t = torch.tensor(acc, device=dev, dtype=torch.float64) / iters
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    names = ["partial kernel (+split merge)", "barrier 0", "peer merge kernel", "barrier 1", "total"]
    print(f"world={world}: " + ", ".join(f"{n} {x*1e3:.1f} us" for n, x in zip(names, t.tolist())))
dist.destroy_process_group()
