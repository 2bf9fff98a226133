"""ctypes binding of ``libpcv_attn.so`` — the C-ABI declared in ``include/pcv_attn.h``.

The structures below mirror the header field by field (tests/test_abi.py checks the sizes and
that every declared symbol is exported).  There is deliberately no fallback: if the shared
library is missing the import of this module's :func:`lib` raises, and every op in
:mod:`perceiver_io_b200.ops` with it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCV_LIB_PATH: developer override (A/B-testing two builds of the library on one GPU box)
LIB_PATH = os.environ.get("PCV_LIB_PATH") or os.path.join(_HERE, "lib", "libpcv_attn.so")

PCV_BF16, PCV_F16, PCV_F32 = 0, 1, 2
PCV_IMPL_AUTO, PCV_IMPL_TCGEN05, PCV_IMPL_SIMT, PCV_IMPL_TCGEN05_PAIR, PCV_IMPL_DECODE = 0, 1, 2, 3, 4
IMPL_BY_NAME = {"auto": PCV_IMPL_AUTO, "tcgen05": PCV_IMPL_TCGEN05, "simt": PCV_IMPL_SIMT, "decode": PCV_IMPL_DECODE,
                "tcgen05_pair": PCV_IMPL_TCGEN05_PAIR}

EXPORTS = (
    "pcv_abi_version",
    "pcv_last_error",
    "pcv_get_device_info",
    "pcv_attn_supported_tcgen05",
    "pcv_attn_workspace_bytes",
    "pcv_attn_fwd",
    "pcv_attn_combine",
    "pcv_attn_combine_peers",
    "pcv_attn_merge_partials",
    "pcv_attn_fwd_sharded_supported",
    "pcv_attn_fwd_sharded",
    "pcv_partial_rescale",
    "pcv_rotary_apply",
    "pcv_kv_append",
    "pcv_kv_project_supported",
    "pcv_ln_stats",
    "pcv_kv_project",
    "pcv_attn_bwd_supported",
    "pcv_attn_bwd_workspace_bytes",
    "pcv_attn_bwd",
    "pcv_attn_fwd_dropout_supported",
    "pcv_attn_fwd_dropout_workspace_bytes",
    "pcv_attn_fwd_dropout",
    "pcv_attn_dropout_mask",
    "pcv_launch_count",
    "pcv_debug_plan",
    "pcv_profile_begin",
    "pcv_profile_end",
    "pcv_debug_read",
    "pcv_debug_trace_read",
)


class AttnParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_stride_b", C.c_int64), ("q_stride_n", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_b", C.c_int64), ("k_stride_m", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_b", C.c_int64), ("v_stride_m", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_b", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_h", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("M", C.c_int32),
        ("dqk", C.c_int32), ("dv", C.c_int32),
        ("scale", C.c_float),
        ("dtype", C.c_int32),
        ("causal", C.c_int32),
        ("m_total", C.c_int32),
        ("m_offset", C.c_int32),
        ("pad_mask", C.c_void_p),
        ("pad_stride_b", C.c_int64),
        ("write_partial", C.c_int32),
        ("part_o", C.c_void_p), ("part_m", C.c_void_p), ("part_l", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("impl", C.c_int32),
        ("reserved", C.c_int32),
    ]


class CombineParams(C.Structure):
    _fields_ = [
        ("part_o", C.c_void_p), ("part_m", C.c_void_p), ("part_l", C.c_void_p), ("out", C.c_void_p),
        ("o_stride_b", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_h", C.c_int64),
        ("num_parts", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("dv", C.c_int32),
        ("dtype", C.c_int32),
    ]


class MergeParams(C.Structure):
    _fields_ = [
        ("part_o", C.c_void_p), ("part_m", C.c_void_p), ("part_l", C.c_void_p),
        ("out_o", C.c_void_p), ("out_m", C.c_void_p), ("out_l", C.c_void_p),
        ("rows", C.c_int64), ("num_parts", C.c_int32), ("dv", C.c_int32),
    ]


PCV_MAX_PEERS = 8


class PeerCombineParams(C.Structure):
    _fields_ = [
        ("part_o", C.c_void_p * PCV_MAX_PEERS), ("part_m", C.c_void_p * PCV_MAX_PEERS),
        ("part_l", C.c_void_p * PCV_MAX_PEERS), ("out", C.c_void_p * PCV_MAX_PEERS),
        ("o_stride_b", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_h", C.c_int64),
        ("row_begin", C.c_int64), ("row_end", C.c_int64),
        ("num_peers", C.c_int32), ("rank", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("dv", C.c_int32),
        ("dtype", C.c_int32), ("reserved", C.c_int32),
    ]


class ShardFuse(C.Structure):
    _fields_ = [
        ("part", C.c_void_p * PCV_MAX_PEERS), ("out", C.c_void_p * PCV_MAX_PEERS), ("flags", C.c_void_p * PCV_MAX_PEERS),
        ("o_stride_b", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_h", C.c_int64),
        ("num_peers", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_uint32), ("reserved", C.c_int32),
    ]


class RescaleParams(C.Structure):
    _fields_ = [
        ("part_o", C.c_void_p), ("part_m", C.c_void_p), ("part_l", C.c_void_p), ("new_m", C.c_void_p),
        ("rows", C.c_int64), ("dv", C.c_int32), ("reserved", C.c_int32),
    ]


class RotaryParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("angles", C.c_void_p),
        ("x_stride_b", C.c_int64), ("x_stride_n", C.c_int64), ("x_stride_h", C.c_int64),
        ("y_stride_b", C.c_int64), ("y_stride_n", C.c_int64), ("y_stride_h", C.c_int64),
        ("a_stride_b", C.c_int64), ("a_stride_n", C.c_int64),
        ("B", C.c_int32), ("n", C.c_int32), ("H", C.c_int32), ("d", C.c_int32),
        ("rotate_dim", C.c_int32),
        ("angle_row0", C.c_int32),
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
    ]


class KvAppendParams(C.Structure):
    _fields_ = [
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("k_new", C.c_void_p), ("v_new", C.c_void_p),
        ("k_dst", C.c_void_p), ("v_dst", C.c_void_p),
        ("kc_stride_b", C.c_int64), ("kc_stride_l", C.c_int64), ("vc_stride_b", C.c_int64), ("vc_stride_l", C.c_int64),
        ("kn_stride_b", C.c_int64), ("kn_stride_l", C.c_int64), ("vn_stride_b", C.c_int64), ("vn_stride_l", C.c_int64),
        ("kd_stride_b", C.c_int64), ("kd_stride_l", C.c_int64), ("vd_stride_b", C.c_int64), ("vd_stride_l", C.c_int64),
        ("B", C.c_int32), ("L_old", C.c_int32), ("n", C.c_int32), ("Ck", C.c_int32), ("Cv", C.c_int32),
        ("dtype", C.c_int32),
    ]


class KvProjParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("col_st", C.c_void_p), ("row_stats", C.c_void_p),
        ("k_out", C.c_void_p), ("v_out", C.c_void_p),
        ("x_stride_row", C.c_int64), ("k_stride_row", C.c_int64), ("v_stride_row", C.c_int64),
        ("rows", C.c_int64),
        ("C", C.c_int32), ("n_k", C.c_int32), ("n_v", C.c_int32),
        ("dtype", C.c_int32), ("cta_group", C.c_int32), ("ln_eps", C.c_float),
    ]


class LnStatsParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("stats", C.c_void_p),
        ("x_stride_row", C.c_int64), ("rows", C.c_int64),
        ("C", C.c_int32), ("eps", C.c_float), ("dtype", C.c_int32), ("reserved", C.c_int32),
    ]


class AttnBwdParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("grad_out", C.c_void_p),
        ("stat_m", C.c_void_p), ("stat_l", C.c_void_p),
        ("grad_q", C.c_void_p), ("grad_k", C.c_void_p), ("grad_v", C.c_void_p),
        ("q_stride_b", C.c_int64), ("q_stride_n", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_stride_b", C.c_int64), ("k_stride_m", C.c_int64), ("k_stride_h", C.c_int64),
        ("v_stride_b", C.c_int64), ("v_stride_m", C.c_int64), ("v_stride_h", C.c_int64),
        ("o_stride_b", C.c_int64), ("o_stride_n", C.c_int64), ("o_stride_h", C.c_int64),
        ("go_stride_b", C.c_int64), ("go_stride_n", C.c_int64), ("go_stride_h", C.c_int64),
        ("gq_stride_b", C.c_int64), ("gq_stride_n", C.c_int64), ("gq_stride_h", C.c_int64),
        ("gk_stride_b", C.c_int64), ("gk_stride_m", C.c_int64), ("gk_stride_h", C.c_int64),
        ("gv_stride_b", C.c_int64), ("gv_stride_m", C.c_int64), ("gv_stride_h", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("M", C.c_int32),
        ("dqk", C.c_int32), ("dv", C.c_int32),
        ("scale", C.c_float), ("dtype", C.c_int32), ("causal", C.c_int32), ("dropout_p", C.c_float),
        ("dropout_seed", C.c_uint64),
        ("pad_mask", C.c_void_p), ("pad_stride_b", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class DeviceInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("sm_major", C.c_int32), ("sm_minor", C.c_int32),
        ("num_sms", C.c_int32), ("smem_optin_bytes", C.c_int32), ("tcgen05_ok", C.c_int32),
    ]


class PcvError(RuntimeError):
    """A libpcv_attn entry point returned a non-zero status."""


_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PcvError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C perceiver_io_b200/csrc`). There is no CPU/PyTorch fallback for the attention path."
            )
        l = C.CDLL(LIB_PATH)
        l.pcv_abi_version.restype = C.c_int
        l.pcv_last_error.restype = C.c_char_p
        l.pcv_launch_count.restype = C.c_uint64
        l.pcv_get_device_info.argtypes = [C.POINTER(DeviceInfo)]
        l.pcv_attn_supported_tcgen05.argtypes = [C.POINTER(AttnParams)]
        l.pcv_attn_workspace_bytes.argtypes = [C.POINTER(AttnParams), C.POINTER(C.c_size_t)]
        l.pcv_attn_fwd.argtypes = [C.POINTER(AttnParams), C.c_void_p]
        l.pcv_attn_combine.argtypes = [C.POINTER(CombineParams), C.c_void_p]
        l.pcv_rotary_apply.argtypes = [C.POINTER(RotaryParams), C.c_void_p]
        l.pcv_partial_rescale.argtypes = [C.POINTER(RescaleParams), C.c_void_p]
        l.pcv_attn_combine_peers.argtypes = [C.POINTER(PeerCombineParams), C.c_void_p]
        l.pcv_attn_combine_peers.restype = C.c_int
        l.pcv_attn_merge_partials.argtypes = [C.POINTER(MergeParams), C.c_void_p]
        l.pcv_attn_merge_partials.restype = C.c_int
        l.pcv_attn_fwd_sharded_supported.argtypes = [C.POINTER(AttnParams)]
        l.pcv_attn_fwd_sharded_supported.restype = C.c_int
        l.pcv_attn_fwd_sharded.argtypes = [C.POINTER(AttnParams), C.POINTER(ShardFuse), C.c_void_p]
        l.pcv_attn_fwd_sharded.restype = C.c_int
        l.pcv_profile_begin.restype = C.c_int
        l.pcv_profile_end.restype = C.c_int
        l.pcv_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        l.pcv_kv_append.argtypes = [C.POINTER(KvAppendParams), C.c_void_p]
        l.pcv_kv_project_supported.argtypes = [C.POINTER(KvProjParams)]
        l.pcv_kv_project_supported.restype = C.c_int
        l.pcv_ln_stats.argtypes = [C.POINTER(LnStatsParams), C.c_void_p]
        l.pcv_ln_stats.restype = C.c_int
        l.pcv_kv_project.argtypes = [C.POINTER(KvProjParams), C.c_void_p]
        l.pcv_kv_project.restype = C.c_int
        l.pcv_attn_bwd_supported.argtypes = [C.POINTER(AttnBwdParams)]
        l.pcv_attn_bwd_supported.restype = C.c_int
        l.pcv_attn_bwd_workspace_bytes.argtypes = [C.POINTER(AttnBwdParams), C.POINTER(C.c_size_t)]
        l.pcv_attn_bwd_workspace_bytes.restype = C.c_int
        l.pcv_attn_bwd.argtypes = [C.POINTER(AttnBwdParams), C.c_void_p]
        l.pcv_attn_bwd.restype = C.c_int
        l.pcv_attn_fwd_dropout_supported.argtypes = [C.POINTER(AttnParams), C.c_float]
        l.pcv_attn_fwd_dropout_supported.restype = C.c_int
        l.pcv_attn_fwd_dropout_workspace_bytes.argtypes = [C.POINTER(AttnParams), C.POINTER(C.c_size_t)]
        l.pcv_attn_fwd_dropout_workspace_bytes.restype = C.c_int
        l.pcv_attn_fwd_dropout.argtypes = [C.POINTER(AttnParams), C.c_void_p, C.c_void_p, C.c_float, C.c_uint64,
                                           C.c_void_p]
        l.pcv_attn_fwd_dropout.restype = C.c_int
        l.pcv_attn_dropout_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                            C.c_uint64, C.c_void_p]
        l.pcv_attn_dropout_mask.restype = C.c_int
        l.pcv_debug_plan.argtypes = [C.c_int32] * 7 + [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]
        l.pcv_debug_plan.restype = C.c_int
        for name in ("pcv_get_device_info", "pcv_attn_supported_tcgen05", "pcv_attn_workspace_bytes",
                     "pcv_attn_fwd", "pcv_attn_combine", "pcv_rotary_apply", "pcv_kv_append",
                     "pcv_partial_rescale"):
            getattr(l, name).restype = C.c_int
        if l.pcv_abi_version() != 1:
            raise PcvError(f"libpcv_attn ABI version {l.pcv_abi_version()} != 1 expected by the Python host")
        _lib = l
        return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().pcv_last_error()
        raise PcvError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def profile_begin() -> None:
    check(lib().pcv_profile_begin(), "pcv_profile_begin")


def profile_end():
    """-> (summed device ms of the attention main kernel, number of launches) since profile_begin()."""
    ms, n = C.c_double(0.0), C.c_int32(0)
    check(lib().pcv_profile_end(C.byref(ms), C.byref(n)), "pcv_profile_end")
    return ms.value, n.value


def debug_read():
    """The tcgen05 kernel's watchdog record (16 words; word 0 != 0 after a barrier-wait timeout)."""
    buf = (C.c_uint32 * 16)()
    l = lib()
    l.pcv_debug_read.restype = C.c_int
    l.pcv_debug_read.argtypes = [C.POINTER(C.c_uint32), C.c_int32]
    l.pcv_debug_read(buf, 16)
    return list(buf)


def debug_plan(B, H, N, M, workers=148, rows_per_unit=256):
    """Host-only: the tcgen05 work plan as (counts dict, list of (cta, b, h, q0, ntile, t0, t1, slot))."""
    counts = (C.c_int32 * 4)()
    lib().pcv_debug_plan(B, H, N, M, workers, rows_per_unit, 128, None, 0, counts)  # sizes only
    n = counts[0]
    segs = (C.c_int32 * (8 * max(n, 1)))()
    check(lib().pcv_debug_plan(B, H, N, M, workers, rows_per_unit, 128, segs, n, counts), "pcv_debug_plan")
    recs = [tuple(segs[8 * i + j] for j in range(8)) for i in range(n)]
    return {"segments": counts[0], "ctas": counts[1], "slots": counts[2], "units": counts[3]}, recs


def launch_count() -> int:
    return int(lib().pcv_launch_count())
