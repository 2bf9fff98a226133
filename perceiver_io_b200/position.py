"""Position helpers of the Perceiver AR path.

Integer path (bit-exact requirement): :func:`positions` restates
/root/reference/perceiver/model/core/position.py:9-17 with the same integer torch ops.
Floating-point path: :class:`RotaryPositionEmbedding` keeps the reference's constructor / attribute
contract (``frq_pos_enc`` (B,1,n,f), ``rotate_dim``, ``right_align``; position.py:20-28) but
``rotate`` runs the CUDA kernel ``pcv_rotary_apply`` instead of materialising cos/sin tensors.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


def positions(b: int, n: int, shift: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """(b, n) int64 absolute positions ``max(arange(n) - shift, 0)`` (left-padding shift per row)."""
    pos = torch.arange(n, device=device).unsqueeze(0).expand(b, n)
    if shift is not None:
        if shift.shape != (b, 1):
            raise ValueError(f"shift must have shape {(1, b)} but has shape {shift.shape}")
        pos = pos - shift
    return torch.clamp(pos, min=0)


class RotaryPositionEmbedding:
    """Holds per-position rotation angles; applies them to the leading channels of each head."""

    def __init__(self, frq_pos_enc: torch.Tensor, right_align: bool = False, inv_freq: Optional[torch.Tensor] = None):
        # (b, n, f) angles, stored broadcastable over heads like the reference does
        self.frq_pos_enc = frq_pos_enc.unsqueeze(1)
        self.rotate_dim = frq_pos_enc.shape[-1]
        self.right_align = right_align
        # optional (not in the reference's signature): the frequency table the angles were built from
        # (FrequencyPositionEncoding.inv_freq).  With it, cached keys can be rotated ONCE, at an absolute position, when
        # they are appended (ops.rotated_cache_keys) instead of re-rotating the whole cache every decode step
        self.inv_freq = inv_freq

    def rotate_rows(self, x: torch.Tensor, num_heads: int) -> torch.Tensor:
        """x: (B, n, H*d) pre-head-split projection output -> rotated copy (fast path used by MHA)."""
        return ops.rotary(x, num_heads, self.frq_pos_enc, self.right_align)

    def rotate(self, t: torch.Tensor) -> torch.Tensor:
        """Reference-signature entry: t is (B, H, n, d) (position.py:30-42)."""
        b, h, n, d = t.shape
        rows = t.permute(0, 2, 1, 3).reshape(b, n, h * d)
        out = self.rotate_rows(rows, h)
        return out.reshape(b, n, h, d).permute(0, 2, 1, 3)


class FrequencyPositionEncoding(nn.Module):
    """angles[b, n, 2i] = angles[b, n, 2i+1] = pos[b, n] * 10000^(-2i/dim) (position.py:53-71)."""

    def __init__(self, dim: int):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))

    def forward(self, abs_pos: torch.Tensor) -> torch.Tensor:
        enc = abs_pos.to(self.inv_freq.dtype).unsqueeze(-1) * self.inv_freq
        return enc.repeat_interleave(2, dim=-1)
