"""ctypes binding of ``_lib/liblis.so`` (C-ABI declared in ``include/lis.h``).

There is deliberately no fallback: if the CUDA library is missing and cannot be built, importing
the scoring entry points raises, and every call on a machine without an sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = Path(os.environ["LIS_LIB"]) if os.environ.get("LIS_LIB") else _PKG / "_lib" / "liblis.so"  # LIS_LIB: A/B builds

LIS_OK = 0
LIS_E_INVALID = -1
LIS_E_CUDA = -2
LIS_E_UNSUPPORTED = -3
LIS_E_NOMEM = -4
LIS_E_NCCL = -5
COMM_ID_BYTES = 128

LIS_BF16 = 0
LIS_F16 = 1
LIS_F32X2 = 2
ROUND_F32 = 0
ROUND_REFERENCE = 1
ROUND_DEFER_SUM = 2
MAX_K = 1024
MTILE = 128
DIM = 128

_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int

# name -> (restype, argtypes); must list every symbol include/lis.h declares (tests check this)
SIGNATURES = {
    "lis_last_error": (C.c_char_p, []),
    "lis_abi_version": (_i32, []),
    "lis_device_supported": (_i32, [_i32]),
    "lis_plan_queries": (_i64, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "lis_maxsim_scores": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32,
                                 _vp, _i64, _vp]),
    "lis_maxsim_pass_plan": (_i32, [_i64, _vp, _i32]),
    "lis_split_f32": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "lis_maxsim_scores_f32x2": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64,
                                       _vp, _i64, _vp]),
    "lis_reduce_segments": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _i64, _vp]),
    "lis_set_tuning": (_i32, [_i32, _i32, _i32, _i32, _i32]),
    "lis_launch_count": (_i64, []),
    "lis_set_ablation": (_i32, [_i32]),
    "lis_set_pass_costs": (_i32, [_vp, _vp]),
    "lis_k1_stats": (_i32, [_vp]),
    "lis_debug_sim_tile": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "lis_debug_sim_pair": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "lis_topk_workspace_bytes": (_i64, [_i64, _i64, _i32]),
    "lis_topk": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp]),
    "lis_merge_topk": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp]),
    "lis_project_normalize": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "lis_index_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i64, _i64]),
    "lis_index_destroy": (None, [_vp]),
    "lis_index_add": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "lis_index_num_pages": (_i64, [_vp]),
    "lis_index_num_rows": (_i64, [_vp]),
    "lis_index_tokens": (_vp, [_vp]),
    "lis_index_tokens_lo": (_vp, [_vp]),
    "lis_index_offsets": (_vp, [_vp]),
    "lis_index_ids": (_vp, [_vp]),
    "lis_index_clamp": (_vp, [_vp]),
    "lis_index_fill_synthetic": (_i32, [_vp, _i64, _vp, _i32, C.c_uint64, _i64, _vp]),
    "lis_index_read_rows": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "lis_index_read_plane": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "lis_index_write_rows": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "lis_index_set_tables": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "lis_index_dtype": (_i32, [_vp]),
    "lis_fill_synthetic_rows": (_i32, [_vp, _i64, _i64, C.c_uint64, _i32, _vp]),
    "lis_index_search": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "lis_index_add_projected": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "lis_index_page_lens": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "lis_comm_unique_id": (_i32, [_vp, _i32]),
    "lis_comm_init": (_i32, [C.POINTER(_vp), _vp, _i32, _i32, _i32]),
    "lis_comm_destroy": (None, [_vp]),
    "lis_comm_rank": (_i32, [_vp]),
    "lis_comm_world": (_i32, [_vp]),
    "lis_nccl_version": (_i32, []),
    "lis_index_search_sharded": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "lis_index_load_rows": (_i32, [_vp, _i32, _i64, _i64, C.c_char_p, _i64, _i32, _vp]),
    "lis_index_save_rows": (_i32, [_vp, _i32, _i64, _i64, C.c_char_p, _i64, _i32, _vp]),
    "lis_index_tombstone": (_i32, [_vp, _i64, _vp]),
    "lis_index_graph_stats": (_i64, [_vp, _vp, _vp]),
    "lis_stream_scores": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp,
                                 _i64, _i64, _i32, _vp]),
    "lis_stream_release": (None, []),
    "lis_memcpy2d_async": (_i32, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load (building first if the sources are newer and nvcc exists) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists() or os.environ.get("LIS_REBUILD") == "1":
        from . import build as _build  # nvcc in-tree build; raises if nvcc is missing

        _build.build(force=os.environ.get("LIS_REBUILD") == "1")
    if not _LIB_PATH.exists():
        raise RuntimeError(f"{_LIB_PATH} is missing and could not be built; there is no fallback path")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.lis_abi_version() != 2:
        raise RuntimeError("liblis.so ABI version mismatch; rebuild with LIS_REBUILD=1")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().lis_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Map a C status code to the exception type the reference's callers expect."""
    if rc >= 0:
        return
    msg = last_error()
    if rc == LIS_E_INVALID:
        raise ValueError(msg)
    if rc == LIS_E_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
