"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): the M-sharded paths — merge fused into the attention
kernel's tail (one launch per rank, several calls in a row), the separate peer-memory kernel and the NCCL protocol —
against the single-GPU result, through torchrun with one rank per GPU (tools/dist_check.py).  The host-side protocol
itself is covered on CPU by tests/test_dist_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs on the box")
def test_sharded_attention_paths_agree_with_one_gpu():
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)   # 2, 4 or 8 ranks
    env = dict(os.environ, PCV_DIST_TIMING="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tools", "dist_check.py")],
                       env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
