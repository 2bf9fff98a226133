"""Hyper-parameter records consumed by the Perceiver AR entry points of the hot path
(``PerceiverAR`` / ``CausalSequenceModel``).  Field names and defaults follow
/root/reference/perceiver/model/core/config.py:64-100 so a reference config's ``asdict`` round-trips;
the encoder/decoder configs of the reference are plain keyword bundles for its task backends and
stay out of scope."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Optional


def _subset(config, base_class, exclude=()):
    names = {f.name for f in dataclasses.fields(base_class)} - set(exclude)
    return {k: v for k, v in dataclasses.asdict(config).items() if k in names}


@dataclass
class PerceiverARConfig:
    num_heads: int = 8
    max_heads_parallel: Optional[int] = None
    num_self_attention_layers: int = 8
    num_self_attention_rotary_layers: int = 1
    self_attention_widening_factor: int = 4
    cross_attention_widening_factor: int = 4
    cross_attention_dropout: float = 0.5
    post_attention_dropout: float = 0.0
    residual_dropout: float = 0.0
    activation_checkpointing: bool = False
    activation_offloading: bool = False

    def base_kwargs(self, exclude=()):
        """Keyword arguments understood by ``PerceiverAR.__init__``."""
        return _subset(self, PerceiverARConfig, exclude)


@dataclass
class CausalSequenceModelConfig(PerceiverARConfig):
    vocab_size: int = 262
    max_seq_len: int = 4096
    max_latents: int = 512
    num_channels: int = 512
    output_norm: bool = False
    output_bias: bool = True
    abs_pos_emb: bool = True
    init_scale: float = 0.02

    @classmethod
    def create(cls, **kwargs):
        known = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in kwargs.items() if k in known})
