"""Cross-attention over key/value inputs that live in HOST memory, pipelined against PCIe.

``CrossAttention.forward`` of the reference takes whatever tensors it is given; when the (B, M, C) input
sits in pinned host memory the copy of M*C elements dominates the step (1.07 GB at the north-star shape,
~20 ms over PCIe Gen5 versus ~3 ms of GPU work).  ``cross_attention_from_host`` splits the key axis into
chunks and overlaps them: while chunk i+1 crosses PCIe on a copy stream, chunk i goes through kv_norm ->
k_proj / v_proj -> the fused attention kernel in *partial* mode (un-normalised numerator, row max,
denominator); the per-chunk states are merged exactly at the end by ``pcv_attn_combine`` — the same algebra
that merges M-shards across GPUs (``dist.py``).  The result equals ``module(x_q, x_kv.cuda())`` up to fp32
re-association in the merge.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .modules import fused_linear, project_kv
from .utils import ModuleOutput


def cross_attention_from_host(module, x_q: torch.Tensor, x_kv_host: torch.Tensor, pad_mask: Optional[torch.Tensor] = None,
                              chunk: int = 8192, device=None, out_host: Optional[torch.Tensor] = None,
                              m_total: Optional[int] = None, m_offset: int = 0, group=None):
    """``module``: a CrossAttention (this package's, or a patched reference one) living on a CUDA device.

    x_q: (1|B, N, D) on the device or in (pinned) host memory; x_kv_host: (B, M, C) in pinned host memory;
    pad_mask: optional (B, M) bool on host or device.  Returns ModuleOutput(last_hidden_state (B, N, F)) on the
    device; if ``out_host`` (pinned) is given the result is also copied into it asynchronously.

    M-sharded use (one process per GPU, ``m_total`` given): ``x_kv_host`` is THIS rank's key shard
    [m_offset, m_offset + M) of ``m_total``; the chunk states are merged locally (``pcv_attn_merge_partials``) into
    the rank's partial state, which is merged across the ranks of ``group`` over peer memory (``dist.PeerMerger``)
    before the replicated ``o_proj``."""
    attn = module.attention
    prm = next(module.parameters())
    device = prm.device if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("cross_attention_from_host needs the module on a CUDA device (no CPU fallback)")
    if attn.causal_attention:
        raise NotImplementedError("streamed host input is implemented for the non-causal encoder cross-attention")
    B, M, _ = x_kv_host.shape
    main = torch.cuda.current_stream(device)
    copy = _copy_stream(device)
    H = attn.num_heads

    with torch.cuda.device(device):
        xq = x_q.to(device, non_blocking=True)
        q = fused_linear(module, "_pcv_q_fold", module.q_norm, attn.q_proj, xq)
        bounds = [(a, min(a + chunk, M)) for a in range(0, M, chunk)]
        G = len(bounds)
        N, dv = q.shape[1], attn.num_v_channels // H
        part_o = torch.empty(G, B, H, N, dv, dtype=torch.float32, device=device)
        part_m = torch.empty(G, B, H, N, dtype=torch.float32, device=device)
        part_l = torch.empty(G, B, H, N, dtype=torch.float32, device=device)
        pad_dev = None if pad_mask is None else pad_mask.to(device, non_blocking=True)

        # Double-buffered device staging of the raw chunk.  The buffers and their "compute done" events persist per
        # device: they are allocated on the MAIN stream (the stream that computes on them; the copy stream waits for
        # that point once), and EVERY reuse of a slot — across calls too — waits for the event recorded after the
        # last kernel that read it, so a copy can never overwrite a chunk that is still being consumed.
        st = _staging(device, B, chunk, x_kv_host.shape[2], prm.dtype, main, copy)
        staged, ready, freed = st["bufs"], st["ready"], st["freed"]

        def issue_copy(i):
            a, b = bounds[i]
            slot = i & 1
            with torch.cuda.stream(copy):
                copy.wait_event(freed[slot])              # compute on the chunk that last used this buffer is done
                view = staged[slot][:, : b - a]
                # one contiguous (rows x C) block per batch row: a strided host slice would be staged through
                # a pageable temporary by torch and serialise the pipeline
                for bi in range(B):
                    view[bi].copy_(x_kv_host[bi, a:b], non_blocking=True)
                ready[slot].record(copy)
            return view

        views = {0: issue_copy(0)}
        for i, (a, b) in enumerate(bounds):
            if i + 1 < G:
                views[i + 1] = issue_copy(i + 1)
            slot = i & 1
            main.wait_event(ready[slot])
            k, v = project_kv(module, views.pop(i))
            ops.attention_partial(q, k, v, H, attn.dp_scale, pad_mask=None if pad_dev is None else pad_dev[:, a:b],
                                  causal=False, m_total=M if m_total is None else m_total, m_offset=m_offset + a,
                                  out=(part_o[i], part_m[i], part_l[i]))
            freed[slot].record(main)
        if m_total is None:
            o = ops.combine_partials(part_o, part_m, part_l, q.dtype)
        else:
            from .dist import PeerMerger

            cdt = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.bfloat16
            pm = PeerMerger.get(B, H, N, dv, cdt, device, group)
            ops.merge_partials(part_o, part_m, part_l, out=(pm.po, pm.pm, pm.pl))
            o = pm.merge()
            o = o if o.dtype == q.dtype else o.to(q.dtype)
        out = fused_linear(attn, "_pcv_o_fold", None, attn.o_proj, o)
        if out_host is not None:
            out_host.copy_(out, non_blocking=True)
    return ModuleOutput(last_hidden_state=out, kv_cache=None)


_streams = {}
_staging_cache = {}


def _staging(device, B, chunk, C, dtype, main, copy):
    key = str(device)
    st = _staging_cache.get(key)
    if st is None or st["shape"] != (B, chunk, C) or st["dtype"] != dtype:
        if st is not None:
            # the old buffers go back to the main stream's pool only after the copy stream is done with them
            for buf in st["bufs"]:
                buf.record_stream(copy)
        bufs = [torch.empty(B, chunk, C, dtype=dtype, device=device) for _ in range(2)]
        allocated = torch.cuda.Event()
        allocated.record(main)
        copy.wait_event(allocated)  # earlier main-stream users of these memory blocks are finished
        st = {"shape": (B, chunk, C), "dtype": dtype, "bufs": bufs,
              "ready": [torch.cuda.Event(), torch.cuda.Event()], "freed": [torch.cuda.Event(), torch.cuda.Event()]}
        _staging_cache[key] = st
    return st


def _copy_stream(device) -> torch.cuda.Stream:
    key = str(device)
    if key not in _streams:
        _streams[key] = torch.cuda.Stream(device=device)
    return _streams[key]
