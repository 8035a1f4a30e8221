"""ctypes binding of liblgcnhs.so (the C ABI declared in include/lgcnhs.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the
product path raises.  (The reference swallows every error in bare ``except:`` blocks and
silently retrains — SURVEY.md §5 — which would hide a missing kernel.)
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "csrc", "liblgcnhs.so"))
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "lgcnhs.h"))

_p = C.c_void_p
_i32, _i64, _f32, _f64, _sz = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/lgcnhs.h one to one
_SIGS = {
    "lgc_abi_version": (C.c_int, []),
    "lgc_last_error_string": (C.c_char_p, []),
    "lgc_launch_count": (_i64, []),
    "lgc_reset_launch_count": (None, []),
    "lgc_csr_max_chunks": (_i64, [_i64]),
    "lgc_csr_build_workspace_bytes": (C.c_int, [_i64, _i64, C.POINTER(_sz)]),
    "lgc_csr_build": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, C.POINTER(_i32), _p, _sz, _p]),
    "lgc_seen_csr_workspace_bytes": (C.c_int, [_i64, C.POINTER(_sz)]),
    "lgc_seen_csr": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, C.POINTER(_i64), _p, _sz, _p]),
    "lgc_unique_u64_workspace_bytes": (C.c_int, [_i64, C.POINTER(_sz)]),
    "lgc_unique_u64": (C.c_int, [_p, _p, _i64, _i32, C.POINTER(_i64), _p, _sz, _p]),
    "lgc_sort_u64_workspace_bytes": (C.c_int, [_i64, C.POINTER(_sz)]),
    "lgc_sort_u64": (C.c_int, [_p, _p, _i64, _i32, _p, _sz, _p]),
    "lgc_spmm_layer": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i64, _i32, _i64, _i64, _p, _i32, _p, _p, _f32, _f32,
                                 _p, _p, _p, _p]),
    "lgc_spmm_layer_bcast": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i64, _i32, _i64, _i64, _p, _i32, _p, _p, _f32,
                                       _f32, C.POINTER(_p), _i32, _p, _p, _p]),
    "lgc_spmm_rows_bcast": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64, _i32, _p, _i64, _i32, _p, _p, _f32, _f32,
                                      C.POINTER(_p), _i32, _p, _p, _p]),
    "lgc_spmm_layer_masked": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i64, _i32, _i64, _i64, _p, _i32, _p, _p, _f32, _f32,
                                        _p, _p, _p, _p, _p]),
    "lgc_spmm_rows_bcast_masked": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64, _i32, _p, _i64, _i32, _p, _p,
                                             _f32, _f32, C.POINTER(_p), _i32, _p, _p, _p, _p]),
    "lgc_propagate_mean_masked": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p,
                                            _p]),
    "lgc_row_mask_batch": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i32, _p]),
    "lgc_peer_barrier": (C.c_int, [_p, C.POINTER(_p), _i32, _i32, _i32, _p]),
    "lgc_peer_barrier_dev": (C.c_int, [_p, C.POINTER(_p), _i32, _i32, _p, _p]),
    "lgc_adam_step_fused": (C.c_int, [_p, C.POINTER(_p), _i32, _p, _p, _p, _p, _i64, _i64, _f32, _f32, _f32, _p, _p]),
    "lgc_zero_rows": (C.c_int, [_p, _p, _i32, _p, _p, _p, _i64, _i64, _p]),
    "lgc_spmm_config": (C.c_int, [_i32]),
    "lgc_spmm_long_row": (C.c_int, [_i32]),
    "lgc_coop_config": (C.c_int, [_i32]),
    "lgc_propagate_mean": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "lgc_propagate_mean_coop": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _p, _p,
                                          _p, _p]),
    "lgc_bpr_scratch_floats": (_i64, [_i64]),
    "lgc_bpr_fwd_bwd": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _i64, _f32, _f32, _p, _p, _p, _p, _p]),
    "lgc_bpr_det_workspace_bytes": (_i64, [_i64, _i32]),
    "lgc_bpr_fwd_bwd_det": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _i64, _f32, _f32, _p, _p, _p, _p, _p, _i64, _p]),
    "lgc_bpr_rows": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _f32, _f32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "lgc_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _p]),
    "lgc_adam_hyper_step": (C.c_int, [_p, _p, _f32, _f32, _p, _p]),
    "lgc_adam_step_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _p, _p]),
    "lgc_score_block": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _p, _p, _f32, _p, _i64, _p]),
    "lgc_topk_rows": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i32, _p, _p, _p]),
    "lgc_score_topk": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _p, _p, _f32, _i32, _p, _i64, _i32, _p, _p, _p]),
    "lgc_score_topk_tc_workspace_bytes": (_i64, [_i64, _i64, _i64, _i32]),
    "lgc_score_topk_tc": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i32, _p, _p, _f32, _i32, _p, _i64, _i32, _p, _p, _p, _i64,
                                    _p]),
    "lgc_negative_sample": (C.c_int, [_p, _p, _i64, _p, _i64, _p, _p, _i64, _i64, _i32, C.c_uint64, _p, _p, _p, _p, _p]),
    "lgc_metrics_scratch_bytes": (_i64, [_i64]),
    "lgc_metrics_topk": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "lgc_score_topk_config": (C.c_int, [_i32]),
    "lgc_mask_from_csr": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "hs_degrees": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "hs_pack_a": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p]),
    "hs_pack_at": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i32, _i32, _p, _p, _i64, _i64, _p]),
    "hs_gemm_planes": (C.c_int, [_i32, _p, _i64, _p, _i64, _i64, _i32, _i64, _i64, _i64, _p, _i64, _p, _p, _f64, _p]),
    "hs_gemm_planes_simt": (C.c_int, [_i32, _p, _i64, _p, _i64, _i64, _i32, _i64, _i64, _i64, _p, _i64, _p, _p,
                                      _f64, _p]),
    "hs_gemm_planes_sym": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _i64, _i64, _p, _i64, _f64, _p]),
    "hs_gemm_planes_sym_bcast": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _i64, _i64, C.POINTER(_p), _i32, _i32, _i32, _i64, _f64,
                                           _p]),
    "hs_resource_topk_scratch_bytes": (_i64, [_i64, _i64, _i32]),
    "hs_resource_topk": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _i64, _i64, _i64, _p, _f64, _p, _i64, _i64, _i32, _p, _p,
                                   _p, _i64, _p]),
    "hs_gemm_config": (C.c_int, [_i32]),
    "hs_gemm_use_cta_pair": (C.c_int, [_i32]),
    "hs_scale_w": (C.c_int, [_p, _i64, _i64, _p, _f64, _p, _i64, _p, _i64, _i64, _i32, _p]),
    "hs_pack_a_u8": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p]),
    "hs_scale_w_u8": (C.c_int, [_p, _i64, _i64, _p, _f64, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _p]),
    "hs_hadamard": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _p]),
    "lgc_probe_gather_threads": (_i64, []),
    "lgc_probe_gather": (C.c_int, [_p, _i64, _i32, _i64, C.c_uint32, _p, C.POINTER(_i64), _p]),
    "lgc_probe_row_store": (C.c_int, [_p, _p, _p, _i64, _i32, _p]),
    "lgc_probe_row_store_persistent": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _p]),
    "lgc_ipc_get_handle": (C.c_int, [_p, _p]),
    "lgc_ipc_open_handle": (C.c_int, [_p, C.POINTER(_p)]),
    "lgc_ipc_close_handle": (C.c_int, [_p]),
}

_LIB = None


class LgcnhsError(RuntimeError):
    pass


def header_symbols() -> list[str]:
    """Every function name include/lgcnhs.h declares (used by the symbol-export test)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:lgc|hs)_[a-z0-9_]+)\s*\(", txt)))


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LgcnhsError(
                f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().lgc_last_error_string().decode(errors="replace")
        raise LgcnhsError(f"liblgcnhs {what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().lgc_launch_count())


def reset_launch_count() -> None:
    lib().lgc_reset_launch_count()
