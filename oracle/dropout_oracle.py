"""CPU restatement (numpy, integer arithmetic) of the counter-based attention-dropout mask of the sm_100a training kernels.

TEST INFRASTRUCTURE ONLY — nothing under perceiver_io_b200/ imports this file (tests/test_abi.py checks that).

What it restates: `drop_bits` / `drop_keep` of perceiver_io_b200/csrc/pcv_attn_bwd.cu, i.e. the generator behind the
dropout that replaces `self.dropout(attn)` of the reference (perceiver/model/core/modules.py:88,161 — nn.Dropout on the
softmax output).  The reference draws its mask from torch's Philox stream; the kernels cannot consume a stored
(B, H, N, M) mask (8.6 GB at the north-star shape), so they regenerate a counter-based one.  Same distribution (each
element dropped independently with probability round(256 p)/256, survivors scaled by the exact inverse keep rate),
different sample — this file pins the sample bit for bit.

    block        = (query >> 1, key >> 1)                      one 32-bit hash per 2 x 2 block
    qword        = (b*H + h) * 0x9E3779B1 + (query >> 1)        (uint32 wrap-around)
    x            = (qword * 0x9E3779B1 ^ seed_lo) ^ ((key >> 1) * 0x85EBCA6B ^ seed_hi)
    x            = hi32(x * 0xD2511F53) ^ lo32(x * 0xD2511F53) ^ 0x9E3779B9
    x            = hi32(x * 0xCD9E8D57) ^ lo32(x * 0xCD9E8D57) ^ 0xBB67AE85
    byte(q, k)   = (x >> 8 * ((query & 1) * 2 + (key & 1))) & 0xff
    keep(q, k)   = byte >= thresh,   thresh = clamp(lround(256 p), 1, 255)   (0 when p == 0: keep everything)
"""
import math

import numpy as np


def drop_threshold(p: float) -> int:
    """Byte threshold and hence the effective drop probability thresh/256 (C: std::lround on the float32 value of p)."""
    if p <= 0.0:
        return 0
    p32 = float(np.float32(p))
    return min(255, max(1, int(math.floor(p32 * 256.0 + 0.5))))


def survivor_scale(p: float) -> float:
    return 256.0 / (256.0 - drop_threshold(p))


def _round(x: np.ndarray, c: int, k: int) -> np.ndarray:
    prod = x.astype(np.uint64) * np.uint64(c)
    return ((prod >> np.uint64(32)).astype(np.uint32) ^ (prod & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            ^ np.uint32(k))


def keep_mask(B: int, H: int, N: int, M: int, p: float, seed: int) -> np.ndarray:
    """(B, H, N, M) bool: True where the element survives."""
    thresh = drop_threshold(p)
    seed_lo, seed_hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    q = np.arange(N, dtype=np.uint32)[:, None]
    k = np.arange(M, dtype=np.uint32)[None, :]
    kside = (k >> np.uint32(1)) * np.uint32(0x85EBCA6B) ^ seed_hi
    shift = ((q & np.uint32(1)) * np.uint32(2) + (k & np.uint32(1))) * np.uint32(8)
    out = np.empty((B * H, N, M), dtype=bool)
    with np.errstate(over="ignore"):
        for bh in range(B * H):
            qword = np.uint32(bh) * np.uint32(0x9E3779B1) + (q >> np.uint32(1))
            x = (qword * np.uint32(0x9E3779B1) ^ seed_lo) ^ kside
            x = _round(x, 0xD2511F53, 0x9E3779B9)
            x = _round(x, 0xCD9E8D57, 0xBB67AE85)
            out[bh] = ((x >> shift) & np.uint32(0xFF)) >= np.uint32(thresh)
    return out.reshape(B, H, N, M)
