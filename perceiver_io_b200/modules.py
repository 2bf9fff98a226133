"""Host-side mirror of the reference's attention building blocks and architectures
(/root/reference/perceiver/model/core/modules.py) with the attention arithmetic replaced by the
sm_100a kernels behind ``include/pcv_attn.h``.

Drop-in contract (level 1): class names, constructor and ``forward`` signatures,
``ModuleOutput`` return type, attribute names read by callers (``q_proj`` ... ``o_proj``,
``dp_scale``, ``num_qk_channels`` ...) and ``state_dict`` keys are those of the reference, so its
Lightning / 🤗 wrappers and checkpoints work on these modules unchanged.  What differs is *how*
``MultiHeadAttention.forward`` computes (reference lines cited inline):

  reference (modules.py:113-170)                      here
  --------------------------------------------------  ------------------------------------------------
  q/k/v/o nn.Linear                                   same (cuBLAS; not on the M-proportional path §8(f)1)
  torch.cat onto the cache                 :117-121   ops.kv_append (one launch for K and V)
  rearrange to (b h n c)                   :123       strides only, never materialised
  q * dp_scale                             :124       folded into the softmax exponent
  rotary on q / k (cos/sin temporaries)    :126-130   ops.rotary (pcv_rotary_apply)
  einsum, 2x masked_fill_, softmax, einsum :146-164   ops.attention (pcv_attn_fwd, online softmax,
                                                      scores never leave the SM)
  max_heads_parallel chunk loop            :144-150   accepted, no effect (nothing to bound)

Inputs must be CUDA tensors; there is no CPU implementation in this package.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .adapter import (InputAdapter, OutputAdapter, QueryProvider, RotarySupport, TiedTokenOutputAdapter,
                      TokenInputAdapterWithRotarySupport, TrainableQueryProvider)
from .config import CausalSequenceModelConfig
from .position import RotaryPositionEmbedding, positions
from .utils import ModuleOutput, Residual, init_parameters

KVCache = Tuple[torch.Tensor, torch.Tensor]


def _rotate_rows(rot, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Apply a rotary embedding object to pre-head-split rows (B, n, H*d).

    Accepts this package's ``RotaryPositionEmbedding`` or any object with the reference's attributes
    (``frq_pos_enc`` (B,1,n,f), ``right_align``) — e.g. one built by reference code."""
    return ops.rotary(x, num_heads, rot.frq_pos_enc, bool(rot.right_align))


class MultiHeadAttention(nn.Module):
    """Multi-head attention with asymmetric query/key lengths, separate qk/v widths, key padding
    mask, right-aligned causal mask, rotary embeddings and a functional KV cache
    (reference modules.py:23-170)."""

    def __init__(
        self,
        num_heads: int,
        num_q_input_channels: int,
        num_kv_input_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        num_output_channels: Optional[int] = None,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        dropout: float = 0.0,
        qkv_bias: bool = True,
        out_bias: bool = True,
    ):
        super().__init__()
        num_qk_channels = num_q_input_channels if num_qk_channels is None else num_qk_channels
        num_v_channels = num_qk_channels if num_v_channels is None else num_v_channels
        num_output_channels = num_q_input_channels if num_output_channels is None else num_output_channels

        if num_qk_channels % num_heads != 0:
            raise ValueError("num_qk_channels must be divisible by num_heads")
        if num_v_channels % num_heads != 0:
            raise ValueError("num_v_channels must be divisible by num_heads")

        self.dp_scale = (num_qk_channels // num_heads) ** -0.5
        self.num_heads = num_heads
        self.num_qk_channels = num_qk_channels
        self.num_v_channels = num_v_channels
        self.causal_attention = causal_attention
        # kept for interface parity; the fused kernel has no (B,h,N,M) tensor to chunk
        self.max_heads_parallel = num_heads if max_heads_parallel is None else max_heads_parallel

        self.q_proj = nn.Linear(num_q_input_channels, num_qk_channels, bias=qkv_bias)
        self.k_proj = nn.Linear(num_kv_input_channels, num_qk_channels, bias=qkv_bias)
        self.v_proj = nn.Linear(num_kv_input_channels, num_v_channels, bias=qkv_bias)
        self.o_proj = nn.Linear(num_v_channels, num_output_channels, bias=out_bias)
        self.dropout = nn.Dropout(dropout)
        self.kernel_impl = "auto"  # "auto" | "tcgen05" | "simt" (testing aid)

    def forward(
        self,
        x_q: torch.Tensor,
        x_kv: torch.Tensor,
        pad_mask: Optional[torch.Tensor] = None,
        rot_pos_emb_q: Optional[RotaryPositionEmbedding] = None,
        rot_pos_emb_k: Optional[RotaryPositionEmbedding] = None,
        kv_cache: Optional[KVCache] = None,
    ):
        """x_q (B|1, N, D), x_kv (B, L, C), pad_mask (B, L_total) bool with True = padding.
        Returns ``ModuleOutput(last_hidden_state=(B, N, F), kv_cache=(k, v) | None)``; cached k/v are
        (B, L_total, channels), un-rotated and pre-head-split exactly like the reference's."""
        q = self.q_proj(x_q)
        k = self.k_proj(x_kv)
        v = self.v_proj(x_kv)
        return attend(self, q, k, v, pad_mask, rot_pos_emb_q, rot_pos_emb_k, kv_cache)


def attend(mha, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pad_mask=None, rot_pos_emb_q=None,
           rot_pos_emb_k=None, kv_cache: Optional[KVCache] = None, min_rows_key: str = "min_rows"):
    """Everything of ``MultiHeadAttention.forward`` after the q/k/v projections (reference modules.py:117-170):
    cache append, rotary, fused attention, ``o_proj``.  ``mha`` is this package's module or a patched reference
    one (only its attributes are used)."""
    # attention-probability dropout (reference :161): fused into the training kernels (ops.attention dropout_p)
    drop_p = float(mha.dropout.p) if mha.training else 0.0
    if kv_cache is not None:
        k, v = ops.kv_append(kv_cache[0], kv_cache[1], k, v)
        kv_cache = (k, v)

    k_att = None
    if (kv_cache is not None and rot_pos_emb_q is not None and rot_pos_emb_k is not None
            and getattr(rot_pos_emb_k, "inv_freq", None) is not None and bool(rot_pos_emb_k.right_align)
            and bool(rot_pos_emb_q.right_align) and not (torch.is_grad_enabled() and (q.requires_grad or k.requires_grad))):
        # decode path: keys are rotated once, when they enter the cache (ops.rotated_cache_keys)
        hit = ops.rotated_cache_keys(k, q, mha.num_heads, rot_pos_emb_k.inv_freq)
        if hit is not None:
            q, k_att = hit
    if k_att is None:
        if rot_pos_emb_q is not None:
            q = _rotate_rows(rot_pos_emb_q, q, mha.num_heads)
        k_att = k if rot_pos_emb_k is None else _rotate_rows(rot_pos_emb_k, k, mha.num_heads)

    o = ops.attention(q, k_att, v, mha.num_heads, mha.dp_scale, pad_mask=pad_mask,
                      causal=mha.causal_attention, impl=getattr(mha, "kernel_impl", "auto"), dropout_p=drop_p)
    o = fused_linear(mha, "_pcv_o_fold", None, mha.o_proj, o, min_rows_key)
    return ModuleOutput(last_hidden_state=o, kv_cache=kv_cache)


#: Policy of the fused K/V producer (LayerNorm + k_proj + v_proj as one tcgen05 GEMM, ``ops.kv_project``).
#: ``min_rows``: below this many key rows the two library GEMMs are used (launch-bound either way).
#: ``min_rows_latent``: threshold of the SELF-attention projections (QKV and o_proj).  In eager mode a small latent array
#: (B*N of a few thousand rows) is bound by host-side dispatch, where ATen's nn.Linear path is leaner than three ctypes
#: calls; under a CUDA graph (``graphs.graph_latent_block`` lowers the threshold while recording) the fused path wins
#: because it launches 4 kernels per layer instead of 7 (tools/latent_stack_bench.py).
kv_producer_config = {"enabled": True, "min_rows": 512, "min_rows_latent": 4096}


def _fold_cache(owner: nn.Module, slot: str, norm: Optional[nn.Module], linears, dtype: torch.dtype):
    """Folded weights of ``norm`` followed by ``linears`` (ops.fold_ln_linear), cached on ``owner`` and rebuilt
    whenever a parameter was modified in place, replaced or moved (data_ptr / ``_version`` of every tensor)."""
    tensors = []
    if norm is not None:
        tensors += [norm.weight, norm.bias]
    for lin in linears:
        tensors += [lin.weight, lin.bias]
    key = (dtype,) + tuple((None if t is None else (t.data_ptr(), t._version)) for t in tensors)
    hit = owner.__dict__.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    w_cat, col_st = ops.fold_ln_linear(None if norm is None else norm.weight, None if norm is None else norm.bias,
                                       [lin.weight for lin in linears], [lin.bias for lin in linears], dtype)
    owner.__dict__[slot] = (key, w_cat, col_st)
    return w_cat, col_st


def _fusable(x: torch.Tensor, linears, norm, min_rows_key: str = "min_rows") -> bool:
    """Inference on bf16/fp16 CUDA rows with parameters in the same dtype: the case the tcgen05 projection kernel
    (ops.kv_project) covers; autograd, autocast, fp32 and tiny inputs stay on LayerNorm + nn.Linear (library GEMMs)."""
    if not kv_producer_config["enabled"] or not x.is_cuda or x.dtype not in (torch.bfloat16, torch.float16):
        return False
    if x.numel() // max(x.shape[-1], 1) < kv_producer_config[min_rows_key] or torch.is_autocast_enabled():
        return False
    if norm is not None and not (isinstance(norm, nn.LayerNorm) and len(norm.normalized_shape) == 1
                                 and norm.normalized_shape[0] == x.shape[-1]):
        return False
    if any(lin.weight.dtype != x.dtype or lin.in_features != x.shape[-1] for lin in linears):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(lin.weight.requires_grad for lin in linears)
                                    or (norm is not None and norm.weight is not None and norm.weight.requires_grad)):
        return False
    return True


def fused_linear(owner: nn.Module, slot: str, norm: Optional[nn.Module], linear: nn.Linear, x: torch.Tensor,
                 min_rows_key: str = "min_rows"):
    """``linear(norm(x))`` (``norm`` may be None) through the tcgen05 projection kernel when it applies — the q_norm ->
    q_proj chain of CrossAttention (reference modules.py:220, :113) and o_proj (:168) — else the library path."""
    n_out = linear.out_features
    if not (_fusable(x, [linear], norm, min_rows_key) and ops.kv_project_supported(x, n_out, 0)):
        return linear(x if norm is None else norm(x))
    affine = norm if (norm is not None and norm.weight is not None) else None
    w_cat, col_st = _fold_cache(owner, slot, affine, [linear], x.dtype)
    y, _ = ops.kv_project(x, w_cat, col_st, n_out, 0, eps=None if norm is None else norm.eps)
    return y


def project_kv(cross_attn, x_kv: torch.Tensor):
    """``k_proj(kv_norm(x_kv)), v_proj(kv_norm(x_kv))`` of a CrossAttention (reference modules.py:226, :114-115).

    Inference on bf16/fp16 CUDA rows goes through the fused producer (one pass over x_kv on the tensor cores,
    LayerNorm folded into the GEMM epilogue, ``pcv_ln_stats`` + ``pcv_kv_project``); everything else — autograd,
    fp32, widths TMA cannot address, tiny inputs — uses LayerNorm + the two ``nn.Linear`` (library GEMMs)."""
    attn = cross_attn.attention
    norm = cross_attn.kv_norm
    n_k, n_v = attn.k_proj.out_features, attn.v_proj.out_features
    if not (_fusable(x_kv, [attn.k_proj, attn.v_proj], norm) and ops.kv_project_supported(x_kv, n_k, n_v)):
        x = norm(x_kv)
        return attn.k_proj(x), attn.v_proj(x)
    w_cat, col_st = _fold_cache(cross_attn, "_pcv_kv_fold", norm if norm.weight is not None else None,
                                [attn.k_proj, attn.v_proj], x_kv.dtype)
    return ops.kv_project(x_kv, w_cat, col_st, n_k, n_v, eps=norm.eps)


def project_qkv(self_attn, x: torch.Tensor):
    """``q_proj(norm(x)), k_proj(norm(x)), v_proj(norm(x))`` of a SelfAttention (reference modules.py:276, :113-115) as
    ONE LayerNorm-folded tcgen05 GEMM over [Wq; Wk; Wv] (``ops.kv_project`` with q as its first output and [k | v] as
    the second: k and v are column ranges of one buffer, the attention kernel takes them by stride).  Returns None when
    the fused path does not apply (autograd, fp32, autocast, tiny inputs): the caller then runs the library path."""
    attn, norm = self_attn.attention, self_attn.norm
    n_q, n_k, n_v = attn.q_proj.out_features, attn.k_proj.out_features, attn.v_proj.out_features
    if not (_fusable(x, [attn.q_proj, attn.k_proj, attn.v_proj], norm, "min_rows_latent")
            and ops.kv_project_supported(x, n_q, n_k + n_v)):
        return None
    w_cat, col_st = _fold_cache(self_attn, "_pcv_qkv_fold", norm if norm.weight is not None else None,
                                [attn.q_proj, attn.k_proj, attn.v_proj], x.dtype)
    q, kv = ops.kv_project(x, w_cat, col_st, n_q, n_k + n_v, eps=norm.eps)
    return q, kv[..., :n_k], kv[..., n_k:]


class CrossAttention(nn.Module):
    """Pre-LayerNorm cross-attention (reference modules.py:173-230)."""

    def __init__(
        self,
        num_heads: int,
        num_q_input_channels: int,
        num_kv_input_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        dropout: float = 0.0,
        qkv_bias: bool = True,
        out_bias: bool = True,
    ):
        super().__init__()
        self.q_norm = nn.LayerNorm(num_q_input_channels)
        self.kv_norm = nn.LayerNorm(num_kv_input_channels)
        self.attention = MultiHeadAttention(
            num_heads=num_heads,
            num_q_input_channels=num_q_input_channels,
            num_kv_input_channels=num_kv_input_channels,
            num_qk_channels=num_qk_channels,
            num_v_channels=num_v_channels,
            max_heads_parallel=max_heads_parallel,
            causal_attention=causal_attention,
            dropout=dropout,
            qkv_bias=qkv_bias,
            out_bias=out_bias,
        )

    def forward(
        self,
        x_q: torch.Tensor,
        x_kv: Optional[torch.Tensor] = None,
        x_kv_prefix: Optional[torch.Tensor] = None,
        pad_mask: Optional[torch.Tensor] = None,
        rot_pos_emb_q: Optional[RotaryPositionEmbedding] = None,
        rot_pos_emb_k: Optional[RotaryPositionEmbedding] = None,
        kv_cache: Optional[KVCache] = None,
    ):
        """With ``x_kv_prefix`` (Perceiver AR) the key/value input is prefix ⧺ query, where the query
        half is normalised by ``q_norm`` and only the prefix by ``kv_norm`` (reference :222-224)."""
        if x_kv is None:
            x_q = self.q_norm(x_q)
            x_kv = torch.cat([self.kv_norm(x_kv_prefix), x_q], dim=1)
            return self.attention(x_q, x_kv, pad_mask=pad_mask, rot_pos_emb_q=rot_pos_emb_q,
                                  rot_pos_emb_k=rot_pos_emb_k, kv_cache=kv_cache)
        q = fused_linear(self, "_pcv_q_fold", self.q_norm, self.attention.q_proj, x_q)
        k, v = project_kv(self, x_kv)
        return attend(self.attention, q, k, v, pad_mask, rot_pos_emb_q, rot_pos_emb_k, kv_cache)


class SelfAttention(nn.Module):
    """Pre-LayerNorm self-attention (reference modules.py:233-278)."""

    def __init__(
        self,
        num_heads: int,
        num_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        dropout: float = 0.0,
        qkv_bias: bool = True,
        out_bias: bool = True,
    ):
        super().__init__()
        self.norm = nn.LayerNorm(num_channels)
        self.attention = MultiHeadAttention(
            num_heads=num_heads,
            num_q_input_channels=num_channels,
            num_kv_input_channels=num_channels,
            num_qk_channels=num_qk_channels,
            num_v_channels=num_v_channels,
            max_heads_parallel=max_heads_parallel,
            causal_attention=causal_attention,
            dropout=dropout,
            qkv_bias=qkv_bias,
            out_bias=out_bias,
        )

    def forward(
        self,
        x: torch.Tensor,
        pad_mask: Optional[torch.Tensor] = None,
        rot_pos_emb: Optional[RotaryPositionEmbedding] = None,
        kv_cache: Optional[KVCache] = None,
    ):
        qkv = project_qkv(self, x)
        if qkv is not None:
            return attend(self.attention, qkv[0], qkv[1], qkv[2], pad_mask, rot_pos_emb, rot_pos_emb, kv_cache,
                          min_rows_key="min_rows_latent")
        x = self.norm(x)
        return self.attention(x, x, pad_mask=pad_mask, rot_pos_emb_q=rot_pos_emb, rot_pos_emb_k=rot_pos_emb,
                              kv_cache=kv_cache)


class AbstractAttentionLayer(nn.Sequential):
    """[attention (optionally residual)] -> [residual MLP]; threads the KV cache through
    (reference modules.py:281-290)."""

    def empty_kv_cache(self, x) -> KVCache:
        shape = (x.shape[0], 0)
        return (torch.empty(*shape, self.num_qk_channels, dtype=x.dtype, device=x.device),
                torch.empty(*shape, self.num_v_channels, dtype=x.dtype, device=x.device))

    def forward(self, *args, kv_cache: Optional[KVCache] = None, **kwargs):
        attended = self[0](*args, kv_cache=kv_cache, **kwargs)
        transformed = self[1](attended.last_hidden_state)
        return ModuleOutput(last_hidden_state=transformed.last_hidden_state, kv_cache=attended.kv_cache)


class CrossAttentionLayer(AbstractAttentionLayer):
    def __init__(
        self,
        num_heads: int,
        num_q_input_channels: int,
        num_kv_input_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        widening_factor: int = 1,
        dropout: float = 0.0,
        residual_dropout: float = 0.0,
        attention_residual: bool = True,
        qkv_bias: bool = True,
        out_bias: bool = True,
        mlp_bias: bool = True,
    ):
        attn = CrossAttention(
            num_heads=num_heads,
            num_q_input_channels=num_q_input_channels,
            num_kv_input_channels=num_kv_input_channels,
            num_qk_channels=num_qk_channels,
            num_v_channels=num_v_channels,
            max_heads_parallel=max_heads_parallel,
            causal_attention=causal_attention,
            dropout=dropout,
            qkv_bias=qkv_bias,
            out_bias=out_bias,
        )
        self.num_qk_channels = attn.attention.num_qk_channels
        self.num_v_channels = attn.attention.num_v_channels
        super().__init__(
            Residual(attn, residual_dropout) if attention_residual else attn,
            Residual(MLP(num_q_input_channels, widening_factor, bias=mlp_bias), residual_dropout),
        )


class SelfAttentionLayer(AbstractAttentionLayer):
    def __init__(
        self,
        num_heads: int,
        num_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        widening_factor: int = 1,
        dropout: float = 0.0,
        residual_dropout: float = 0.0,
        qkv_bias: bool = True,
        out_bias: bool = True,
        mlp_bias: bool = True,
    ):
        attn = SelfAttention(
            num_heads=num_heads,
            num_channels=num_channels,
            num_qk_channels=num_qk_channels,
            num_v_channels=num_v_channels,
            max_heads_parallel=max_heads_parallel,
            causal_attention=causal_attention,
            dropout=dropout,
            qkv_bias=qkv_bias,
            out_bias=out_bias,
        )
        self.num_qk_channels = attn.attention.num_qk_channels
        self.num_v_channels = attn.attention.num_v_channels
        super().__init__(
            Residual(attn, residual_dropout),
            Residual(MLP(num_channels, widening_factor, bias=mlp_bias), residual_dropout),
        )


class SelfAttentionBlock(nn.Sequential):
    """Stack of self-attention layers; rotary only in the first ``num_rotary_layers`` (all if -1);
    one KV-cache pair per layer, ``[]`` meaning "initialise" (reference modules.py:370-441)."""

    def __init__(
        self,
        num_layers: int,
        num_heads: int,
        num_channels: int,
        num_qk_channels: Optional[int] = None,
        num_v_channels: Optional[int] = None,
        num_rotary_layers: int = 1,
        max_heads_parallel: Optional[int] = None,
        causal_attention: bool = False,
        widening_factor: int = 1,
        dropout: float = 0.0,
        residual_dropout: float = 0.0,
        activation_checkpointing: bool = False,
        activation_offloading: bool = False,
        qkv_bias: bool = True,
        out_bias: bool = True,
        mlp_bias: bool = True,
    ):
        # activation_checkpointing / activation_offloading: accepted for signature parity.  The fused
        # kernel never stores the (B,h,N,M) probabilities, which is what checkpointing was saving.
        super().__init__(*[
            SelfAttentionLayer(
                num_heads=num_heads,
                num_channels=num_channels,
                num_qk_channels=num_qk_channels,
                num_v_channels=num_v_channels,
                max_heads_parallel=max_heads_parallel,
                causal_attention=causal_attention,
                widening_factor=widening_factor,
                dropout=dropout,
                residual_dropout=residual_dropout,
                qkv_bias=qkv_bias,
                out_bias=out_bias,
                mlp_bias=mlp_bias,
            )
            for _ in range(num_layers)
        ])
        self.num_rotary_layers = num_rotary_layers

    def forward(
        self,
        x: torch.Tensor,
        pad_mask: Optional[torch.Tensor] = None,
        rot_pos_emb: Optional[RotaryPositionEmbedding] = None,
        kv_cache: Optional[List[KVCache]] = None,
    ):
        new_cache: Optional[List[KVCache]] = None
        if kv_cache is not None:
            if len(kv_cache) == 0:
                kv_cache = [layer.empty_kv_cache(x) for layer in self]
            new_cache = []

        for idx, layer in enumerate(self):
            use_rot = self.num_rotary_layers == -1 or idx < self.num_rotary_layers
            out = layer(x, pad_mask=pad_mask, rot_pos_emb=rot_pos_emb if use_rot else None,
                        kv_cache=None if kv_cache is None else kv_cache[idx])
            x = out.last_hidden_state
            if new_cache is not None:
                new_cache.append(out.kv_cache)

        return ModuleOutput(last_hidden_state=x, kv_cache=new_cache)


class MLP(nn.Sequential):
    """LayerNorm -> Linear -> GELU -> Linear (reference modules.py:444-454); stays on cuBLAS."""

    def __init__(self, num_channels: int, widening_factor: int, bias: bool = True):
        super().__init__(
            nn.LayerNorm(num_channels),
            nn.Linear(num_channels, widening_factor * num_channels, bias=bias),
            nn.GELU(),
            nn.Linear(widening_factor * num_channels, num_channels, bias=bias),
        )

    def forward(self, x):
        return ModuleOutput(last_hidden_state=super().forward(x))


class PerceiverEncoder(nn.Module):
    """Latents (1, N, D) cross-attend to adapted inputs (B, M, C), then run through self-attention
    blocks; optional repeated cross-attention and weight sharing (reference modules.py:457-607)."""

    def __init__(
        self,
        input_adapter: InputAdapter,
        num_latents: int,
        num_latent_channels: int,
        num_cross_attention_heads: int = 4,
        num_cross_attention_qk_channels: Optional[int] = None,
        num_cross_attention_v_channels: Optional[int] = None,
        num_cross_attention_layers: int = 1,
        first_cross_attention_layer_shared: bool = False,
        cross_attention_widening_factor: int = 1,
        num_self_attention_heads: int = 4,
        num_self_attention_qk_channels: Optional[int] = None,
        num_self_attention_v_channels: Optional[int] = None,
        num_self_attention_layers_per_block: int = 6,
        num_self_attention_blocks: int = 1,
        first_self_attention_block_shared: bool = True,
        self_attention_widening_factor: int = 1,
        dropout: float = 0.0,
        residual_dropout: float = 0.0,
        init_scale: float = 0.02,
        activation_checkpointing: bool = False,
        activation_offloading: bool = False,
    ):
        super().__init__()
        self.latent_provider = TrainableQueryProvider(num_latents, num_latent_channels, init_scale=init_scale)
        self.input_adapter = input_adapter

        if num_cross_attention_layers <= 0:
            raise ValueError("num_cross_attention_layers must be > 0")
        if num_self_attention_blocks <= 0:
            raise ValueError("num_self_attention_blocks must be > 0")
        if num_cross_attention_layers > num_self_attention_blocks:
            raise ValueError("num_cross_attention_layers must be <= num_self_attention_blocks")

        self.num_cross_attention_layers = num_cross_attention_layers
        self.num_self_attention_blocks = num_self_attention_blocks
        self.first_cross_attention_layer_shared = first_cross_attention_layer_shared
        self.first_self_attention_block_shared = first_self_attention_block_shared

        def make_cross_attn():
            return CrossAttentionLayer(
                num_heads=num_cross_attention_heads,
                num_q_input_channels=num_latent_channels,
                num_kv_input_channels=input_adapter.num_input_channels,
                num_qk_channels=num_cross_attention_qk_channels,
                num_v_channels=num_cross_attention_v_channels,
                widening_factor=cross_attention_widening_factor,
                dropout=dropout,
                residual_dropout=residual_dropout,
            )

        def make_self_attn():
            return SelfAttentionBlock(
                num_layers=num_self_attention_layers_per_block,
                num_heads=num_self_attention_heads,
                num_channels=num_latent_channels,
                num_qk_channels=num_self_attention_qk_channels,
                num_v_channels=num_self_attention_v_channels,
                widening_factor=self_attention_widening_factor,
                dropout=dropout,
                residual_dropout=residual_dropout,
                activation_checkpointing=activation_checkpointing,
                activation_offloading=activation_offloading,
            )

        self.cross_attn_1 = make_cross_attn()
        self.self_attn_1 = make_self_attn()
        if self.extra_cross_attention_layer:
            self.cross_attn_n = make_cross_attn()
        if self.extra_self_attention_block:
            self.self_attn_n = make_self_attn()

        self._init_parameters(init_scale)

    def _init_parameters(self, init_scale: float):
        with torch.no_grad():
            init_parameters(self, init_scale)

    @property
    def extra_cross_attention_layer(self):
        return self.num_cross_attention_layers > 1 and not self.first_cross_attention_layer_shared

    @property
    def extra_self_attention_block(self):
        return self.num_self_attention_blocks > 1 and not self.first_self_attention_block_shared

    def forward(self, x, pad_mask=None, return_adapted_input=False):
        x_adapted = self.input_adapter(x)
        x_latent = self.latent_provider()  # (1, N, D): broadcast by the kernel, never expanded

        x_latent = self.cross_attn_1(x_latent, x_adapted, pad_mask=pad_mask).last_hidden_state
        x_latent = self.self_attn_1(x_latent).last_hidden_state

        later_cross = self.cross_attn_n if self.extra_cross_attention_layer else self.cross_attn_1
        later_self = self.self_attn_n if self.extra_self_attention_block else self.self_attn_1
        for block in range(1, self.num_self_attention_blocks):
            if block < self.num_cross_attention_layers:
                x_latent = later_cross(x_latent, x_adapted, pad_mask=pad_mask).last_hidden_state
            x_latent = later_self(x_latent).last_hidden_state

        return (x_latent, x_adapted) if return_adapted_input else x_latent


class PerceiverDecoder(nn.Module):
    """Output queries cross-attend to the latents (reverse asymmetry: many queries, few keys);
    reference modules.py:610-675."""

    def __init__(
        self,
        output_adapter: OutputAdapter,
        output_query_provider: QueryProvider,
        num_latent_channels: int,
        num_cross_attention_heads: int = 4,
        num_cross_attention_qk_channels: Optional[int] = None,
        num_cross_attention_v_channels: Optional[int] = None,
        cross_attention_widening_factor: int = 1,
        cross_attention_residual: bool = True,
        dropout: float = 0.0,
        init_scale: float = 0.02,
        activation_checkpointing: bool = False,
        activation_offloading: bool = False,
    ):
        super().__init__()
        self.output_query_provider = output_query_provider
        self.output_adapter = output_adapter
        self.cross_attn = CrossAttentionLayer(
            num_heads=num_cross_attention_heads,
            num_q_input_channels=output_query_provider.num_query_channels,
            num_kv_input_channels=num_latent_channels,
            num_qk_channels=num_cross_attention_qk_channels,
            num_v_channels=num_cross_attention_v_channels,
            widening_factor=cross_attention_widening_factor,
            attention_residual=cross_attention_residual,
            dropout=dropout,
        )
        self._init_parameters(init_scale)

    def _init_parameters(self, init_scale: float):
        with torch.no_grad():
            init_parameters(self, init_scale)

    def forward(self, x_latent, x_adapted=None, **kwargs):
        output_query = self.output_query_provider(x_adapted)
        decoded = self.cross_attn(output_query, x_latent).last_hidden_state
        return self.output_adapter(decoded, **kwargs)


class PerceiverIO(nn.Sequential):
    def __init__(self, encoder: PerceiverEncoder, decoder: PerceiverDecoder):
        super().__init__(encoder, decoder)

    @property
    def encoder(self):
        return self[0]

    @property
    def decoder(self):
        return self[1]


class PerceiverAR(nn.Module):
    """Perceiver AR: the last ``n - prefix_len`` positions are latents that cross-attend causally to
    prefix ⧺ latents, followed by causal latent self-attention; left padding, prefix dropout,
    right-aligned rotary embeddings and KV caching as in reference modules.py:691-871."""

    def __init__(
        self,
        input_adapter: RotarySupport,
        num_heads: int = 8,
        max_heads_parallel: Optional[int] = None,
        num_self_attention_layers: int = 6,
        num_self_attention_rotary_layers: int = 1,
        self_attention_widening_factor: int = 4,
        cross_attention_widening_factor: int = 4,
        cross_attention_dropout: float = 0.5,
        post_attention_dropout: float = 0.0,
        residual_dropout: float = 0.0,
        activation_checkpointing: bool = False,
        activation_offloading: bool = False,
    ):
        super().__init__()
        channels = input_adapter.num_input_channels
        self.input_adapter = input_adapter
        self.cross_attention_dropout = cross_attention_dropout
        self.cross_attention = CrossAttentionLayer(
            num_heads=num_heads,
            num_q_input_channels=channels,
            num_kv_input_channels=channels,
            max_heads_parallel=max_heads_parallel,
            causal_attention=True,
            widening_factor=cross_attention_widening_factor,
            dropout=post_attention_dropout,
            residual_dropout=residual_dropout,
            qkv_bias=False,
            out_bias=True,
            mlp_bias=False,
        )
        self.self_attention = SelfAttentionBlock(
            num_layers=num_self_attention_layers,
            num_heads=num_heads,
            num_channels=channels,
            causal_attention=True,
            widening_factor=self_attention_widening_factor,
            dropout=post_attention_dropout,
            residual_dropout=residual_dropout,
            num_rotary_layers=num_self_attention_rotary_layers,
            activation_checkpointing=activation_checkpointing,
            activation_offloading=activation_offloading,
            qkv_bias=False,
            out_bias=False,
            mlp_bias=False,
        )

    def forward(
        self,
        x: torch.Tensor,
        prefix_len: int,
        pad_mask: Optional[torch.Tensor] = None,
        kv_cache: Optional[List[KVCache]] = None,
    ):
        # ---- integer path (must match the reference bit for bit; modules.py:776-807) -------------
        shift = None if pad_mask is None else pad_mask.sum(dim=1, keepdim=True)  # x is left-padded
        cache_active = kv_cache is not None and len(kv_cache) > 0
        b = x.shape[0]
        n = x.shape[1] + (kv_cache[0][0].shape[1] if cache_active else 0)
        if not 0 <= prefix_len < n:
            raise ValueError(f"prefix_len ({prefix_len}) out of valid range [0..{n})")

        x, frq_pos_enc = self.input_adapter(x, abs_pos=positions(b, n, shift=shift, device=x.device))

        if cache_active:
            x_latent, x_prefix = x, x[:, :0]
        else:
            x_latent, x_prefix = x[:, prefix_len:], x[:, :prefix_len]

        frq_latent, frq_prefix = frq_pos_enc[:, prefix_len:], frq_pos_enc[:, :prefix_len]
        if pad_mask is not None:
            pad_latent, pad_prefix = pad_mask[:, prefix_len:], pad_mask[:, :prefix_len]

        # ---- training-time prefix dropout: keep a random subset of prefix positions (:809-830) ----
        if self.training and prefix_len > 0 and self.cross_attention_dropout > 0.0:
            if kv_cache is not None:
                raise ValueError("cross-attention dropout not supported with caching")
            rand = torch.rand(b, prefix_len, device=x.device)
            keep = prefix_len - int(prefix_len * self.cross_attention_dropout)
            keep_idx = rand.topk(keep, dim=-1).indices
            keep_mask = torch.zeros_like(rand, dtype=torch.bool).scatter_(dim=1, index=keep_idx, value=1)
            x_prefix = x_prefix[keep_mask].reshape(b, keep, x_prefix.shape[-1])
            frq_prefix = frq_prefix[keep_mask].reshape(b, keep, frq_prefix.shape[-1])
            if pad_mask is not None:
                pad_prefix = pad_prefix[keep_mask].reshape(b, keep)

        frq_keys = torch.cat([frq_prefix, frq_latent], dim=1)
        if pad_mask is not None:
            pad_mask = torch.cat([pad_prefix, pad_latent], dim=1)

        # ---- cache routing (:838-848) -------------------------------------------------------------
        if kv_cache is None:
            ca_cache, sa_cache, new_cache = None, None, None
        elif len(kv_cache) == 0:
            ca_cache, sa_cache, new_cache = self.cross_attention.empty_kv_cache(x_latent), [], []
        else:
            ca_cache, sa_cache, new_cache = kv_cache[0], list(kv_cache[1:]), []

        # frequency table of the adapter: lets cached decoding rotate new keys only (ops.rotated_cache_keys)
        inv_freq = getattr(getattr(self.input_adapter, "frq_pos_encoding", None), "inv_freq", None)
        ca_out = self.cross_attention(
            x_latent,
            x_kv_prefix=x_prefix,
            pad_mask=pad_mask,
            rot_pos_emb_q=RotaryPositionEmbedding(frq_latent, right_align=True, inv_freq=inv_freq),
            rot_pos_emb_k=RotaryPositionEmbedding(frq_keys, right_align=True, inv_freq=inv_freq),
            kv_cache=ca_cache,
        )
        if new_cache is not None:
            new_cache.append(ca_out.kv_cache)

        sa_out = self.self_attention(
            ca_out.last_hidden_state,
            rot_pos_emb=RotaryPositionEmbedding(frq_latent, right_align=True, inv_freq=inv_freq),
            kv_cache=sa_cache,
        )
        if new_cache is not None:
            new_cache.extend(sa_out.kv_cache)

        return ModuleOutput(last_hidden_state=sa_out.last_hidden_state, kv_cache=new_cache)


class CausalSequenceModel(PerceiverAR):
    """Perceiver AR + token adapters + tied-embedding logits (reference modules.py:874-930)."""

    def __init__(self, config: CausalSequenceModelConfig):
        rotated = config.num_channels // config.num_heads
        if config.abs_pos_emb:
            rotated //= 2  # rotary on the first half of each head's channels only
        input_adapter = TokenInputAdapterWithRotarySupport(
            rotated_channels_per_head=rotated,
            vocab_size=config.vocab_size,
            max_seq_len=config.max_seq_len,
            num_input_channels=config.num_channels,
            abs_pos_emb=config.abs_pos_emb,
        )
        super().__init__(input_adapter=input_adapter, **config.base_kwargs())
        self.config = config
        if config.output_norm:
            self.out_norm = nn.LayerNorm(config.num_channels)
        self.output_adapter = TiedTokenOutputAdapter(vocab_size=config.vocab_size, emb_bias=config.output_bias)
        self._init_parameters(config.init_scale)

    def _init_parameters(self, init_scale: float):
        with torch.no_grad():
            init_parameters(self, init_scale)

    @property
    def max_seq_len(self):
        return self.input_adapter.max_seq_len

    @property
    def max_latents(self):
        return self.config.max_latents

    @property
    def max_prefix_len(self):
        return self.max_seq_len - self.max_latents

    def forward(
        self,
        x: torch.Tensor,
        prefix_len: int,
        pad_mask: Optional[torch.Tensor] = None,
        kv_cache: Optional[List[KVCache]] = None,
    ):
        if prefix_len > self.max_prefix_len:
            raise ValueError(f"prefix_len ({prefix_len}) exceeds max_prefix_len ({self.max_prefix_len})")
        output = super().forward(x, prefix_len=prefix_len, pad_mask=pad_mask, kv_cache=kv_cache)
        if self.config.output_norm:
            output.last_hidden_state = self.out_norm(output.last_hidden_state)
        output.logits = self.output_adapter(output.last_hidden_state, txt_embedding=self.input_adapter.txt_embedding)
        return output
