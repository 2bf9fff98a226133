"""Return/cache types and the residual wrapper of the latent-attention path.

Behavioural mirror of /root/reference/perceiver/model/core/utils.py:6-47 (``ModuleOutput``,
``Residual``, ``init_parameters``, ``freeze``); these are part of the drop-in boundary
(SURVEY.md §8(a) A4/A11): callers read ``output.last_hidden_state`` / ``output.kv_cache`` and the
residual adds the *un-normed* first positional argument after dropout.
"""
from __future__ import annotations

from collections import OrderedDict

import torch.nn as nn


class ModuleOutput(OrderedDict):
    """Ordered dict whose items are also attributes (``out.kv_cache is out["kv_cache"]``)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError("No such attribute: " + key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError:
            raise AttributeError("No such attribute: " + key) from None


class Residual(nn.Module):
    """``y = dropout(module(x, ...).last_hidden_state) + x`` with x the first positional input.

    x may have batch 1 against a batch-B module output (encoder latents): the add broadcasts."""

    def __init__(self, module: nn.Module, dropout: float = 0.0):
        super().__init__()
        self.module = module
        self.dropout = nn.Dropout(dropout)

    def forward(self, *args, **kwargs):
        result = self.module(*args, **kwargs)
        result.last_hidden_state = self.dropout(result.last_hidden_state) + args[0]
        return result


def init_parameters(module: nn.Module, init_scale: float) -> None:
    """N(0, init_scale) for Linear/Embedding weights, zero Linear biases (reference utils.py:35-42)."""
    for sub in module.modules():
        if isinstance(sub, nn.Linear):
            sub.weight.data.normal_(mean=0.0, std=init_scale)
            if sub.bias is not None:
                sub.bias.data.zero_()
        elif isinstance(sub, nn.Embedding):
            sub.weight.data.normal_(mean=0.0, std=init_scale)


def freeze(module: nn.Module) -> None:
    for prm in module.parameters():
        prm.requires_grad = False
