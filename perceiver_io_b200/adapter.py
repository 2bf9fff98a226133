"""Adapters / query providers the hot path's callers need (out of scope as
compute — embedding lookups and one GEMM stay in PyTorch — but their interfaces feed the path:
``TrainableQueryProvider`` returns the batch-1 latent array the kernel broadcasts, and
``RotarySupport`` supplies the angle tensor the rotary kernel consumes).

Interface mirror of /root/reference/perceiver/model/core/adapter.py (same class names, constructor
arguments, properties and parameter names so reference checkpoints load unchanged).
"""
from __future__ import annotations

import torch
from torch import nn

from .position import FrequencyPositionEncoding, positions


class InputAdapter(nn.Module):
    """Maps task input to the (B, M, C) key/value input of the encoder; C = ``num_input_channels``."""

    def __init__(self, num_input_channels: int, *args, **kwargs):
        super().__init__()
        self._num_input_channels = num_input_channels

    @property
    def num_input_channels(self) -> int:
        return self._num_input_channels


class RotarySupport(InputAdapter):
    """Mixin: ``forward`` additionally returns rotation angles for the (shifted) positions."""

    def __init__(self, rotated_channels_per_head: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.frq_pos_encoding = FrequencyPositionEncoding(dim=rotated_channels_per_head)

    def forward(self, x, abs_pos=None):
        if abs_pos is None:
            abs_pos = positions(*x.shape, device=x.device)
        return super().forward(x, abs_pos), self.frq_pos_encoding(abs_pos)


class OutputAdapter(nn.Module):
    """Maps decoder cross-attention output (B, O, F) to task output."""


class ClassificationOutputAdapter(OutputAdapter):
    def __init__(self, num_classes: int, num_output_query_channels: int):
        super().__init__()
        self.linear = nn.Linear(num_output_query_channels, num_classes)

    def forward(self, x):
        return self.linear(x).squeeze(dim=1)


class QueryProvider:
    """Source of a cross-attention query array."""

    @property
    def num_query_channels(self):
        raise NotImplementedError()

    def __call__(self, x=None):
        raise NotImplementedError()


class TrainableQueryProvider(nn.Module, QueryProvider):
    """Learned (num_queries, C) array returned with a leading batch dimension of ONE — the latent
    array of the encoder / the output queries of most decoders.  The attention kernel broadcasts it
    over the batch with a zero batch stride instead of expanding it."""

    def __init__(self, num_queries: int, num_query_channels: int, init_scale: float = 0.02):
        super().__init__()
        self._query = nn.Parameter(torch.empty(num_queries, num_query_channels))
        with torch.no_grad():
            self._query.normal_(0.0, init_scale)

    @property
    def num_query_channels(self):
        return self._query.shape[-1]

    def forward(self, x=None):
        return self._query.unsqueeze(0)


class TokenInputAdapter(InputAdapter):
    def __init__(self, vocab_size: int, max_seq_len: int, num_input_channels: int, abs_pos_emb: bool = True):
        super().__init__(num_input_channels)
        self._max_seq_len = max_seq_len
        self._abs_pos_emb = abs_pos_emb
        self.txt_embedding = nn.Embedding(vocab_size, num_input_channels)
        if abs_pos_emb:
            self.pos_embedding = nn.Embedding(max_seq_len, num_input_channels)

    @property
    def vocab_size(self):
        return self.txt_embedding.num_embeddings

    @property
    def max_seq_len(self):
        return self._max_seq_len

    def forward(self, x, abs_pos=None):
        emb = self.txt_embedding(x)
        if not self._abs_pos_emb:
            return emb
        if abs_pos is None:
            abs_pos = positions(*x.shape, device=x.device)
        elif x.shape[1] < abs_pos.shape[1]:
            abs_pos = abs_pos[:, -x.shape[1]:]  # cached decoding: right-most position codes
        return emb + self.pos_embedding(abs_pos)


class TokenInputAdapterWithRotarySupport(RotarySupport, TokenInputAdapter):
    def __init__(self, rotated_channels_per_head: int, vocab_size: int, max_seq_len: int,
                 num_input_channels: int, abs_pos_emb: bool):
        super().__init__(
            rotated_channels_per_head=rotated_channels_per_head,
            vocab_size=vocab_size,
            max_seq_len=max_seq_len,
            num_input_channels=num_input_channels,
            abs_pos_emb=abs_pos_emb,
        )

    def forward(self, x, abs_pos=None):
        return super().forward(x, abs_pos)


class TiedTokenOutputAdapter(OutputAdapter):
    """logits = h @ E^T (+ bias) with E the input embedding matrix."""

    def __init__(self, vocab_size: int, emb_bias: bool = True):
        super().__init__()
        self._emb_bias = emb_bias
        if emb_bias:
            self.bias = nn.Parameter(torch.zeros(vocab_size))

    def forward(self, x, txt_embedding: nn.Embedding):
        logits = torch.matmul(x, txt_embedding.weight.T)
        return logits + self.bias if self._emb_bias else logits
