"""ctypes binding of libkemr.so (include/kemr.h).  There is no fallback: if the CUDA library
is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KEMR_LIB selects another build of the same library (the debug build with the per-role cycle counters, tools/)
LIB_PATH = os.environ.get("KEMR_LIB") or os.path.join(_HERE, "libkemr.so")

KEMR_OK = 0
PATH_AUTO, PATH_WARP, PATH_MMA = 0, 1, 2
FLAG_UNCERTIFIED, FLAG_OVERFLOW = 1, 2
ABI_VERSION = 4

EXPORTS = (
    "kemr_last_error", "kemr_abi_version", "kemr_device_info", "kemr_quantize_rows", "kemr_synth_rows",
    "kemr_workspace_bytes", "kemr_scan_topk", "kemr_score_pairs", "kemr_rank_count", "kemr_score_matrix",
    "kemr_matrix_rank", "kemr_matrix_topk", "kemr_matrix_fuse", "kemr_metrics_reduce",
    "kemr_metrics_reduce_host", "kemr_merge_topk", "kemr_index_create", "kemr_index_destroy",
    "kemr_index_search_host", "kemr_set_scan_done_event", "kemr_scan_plan",
    "kemr_scan_topk_gated", "kemr_rank_count_gated", "kemr_score_pairs_gated", "kemr_gate_linear",
    "kemr_hits_workspace_bytes", "kemr_hits_build_csr", "kemr_idmap_create", "kemr_idmap_destroy", "kemr_idmap_lookup",
    "kemr_store_write", "kemr_store_info", "kemr_store_load", "kemr_debug_mma_plan", "kemr_set_phase_stamps",
    "kemr_peer_create", "kemr_peer_connect", "kemr_peer_connect_pointers", "kemr_peer_local_buffer", "kemr_peer_destroy",
    "kemr_peer_begin", "kemr_peer_merge", "kemr_merge_topk_strided",
    "kemr_hits_filter_csr", "kemr_hits_target_bonus", "kemr_matrix_mlp2", "kemr_infonce_rows", "kemr_row_norm_max", "kemr_debug_select_stamps", "kemr_index_search_host_bf16", "kemr_peer_gather",
    "kemr_index_share", "kemr_index_submit_host", "kemr_index_wait",
)


class KemrError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    p, i32, i64, u64, f32, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_size_t
    lib.kemr_last_error.restype = C.c_char_p
    lib.kemr_last_error.argtypes = []
    lib.kemr_abi_version.restype = i32
    lib.kemr_abi_version.argtypes = []
    lib.kemr_device_info.argtypes = [C.POINTER(i32)] * 4
    lib.kemr_quantize_rows.argtypes = [p, p, i64, i32, i32, p]
    lib.kemr_row_norm_max.argtypes = [p, i64, i32, p, p]
    lib.kemr_synth_rows.argtypes = [p, i64, i32, u64, i64, p]
    lib.kemr_scan_plan.argtypes = [i32, i64, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.kemr_workspace_bytes.restype = sz
    lib.kemr_workspace_bytes.argtypes = [i32, i64, i32, i32, i64]
    lib.kemr_scan_topk.argtypes = [p, i32, p, p, i64, i32, f64, f64, f64, p, p, p, i64, i32, i32, f64, i64,
                                   p, p, p, p, p, sz, i32, p]
    lib.kemr_score_pairs.argtypes = [p, p, p, i64, i32, f64, f64, f64, p, p, p, i64, p, p]
    lib.kemr_scan_topk_gated.argtypes = [p, i32, p, p, i64, i32, p, p, f64, p, p, p, i64, i32, i32, f64, i64,
                                         p, p, p, p, p, sz, i32, p]
    lib.kemr_rank_count_gated.argtypes = [p, i32, p, p, i64, i32, p, p, f64, p, p, p, p, p, f64, i64,
                                          p, p, p, sz, i32, p]
    lib.kemr_score_pairs_gated.argtypes = [p, p, p, i64, i32, p, p, f64, p, p, p, i64, p, p]
    lib.kemr_gate_linear.argtypes = [p, i32, i32, p, f32, p, p, p]
    lib.kemr_hits_workspace_bytes.restype = sz
    lib.kemr_hits_workspace_bytes.argtypes = [i32]
    lib.kemr_hits_build_csr.argtypes = [p, p, p, i32, i64, i64, i32, p, p, p, p, p, sz, p]
    lib.kemr_hits_filter_csr.argtypes = [p, p, p, p, i32, i64, i64, p, p, p, p, p, sz, p]
    lib.kemr_hits_target_bonus.argtypes = [p, p, p, i32, p, p, p]
    lib.kemr_idmap_create.argtypes = [p, p, i64, C.POINTER(p)]
    lib.kemr_idmap_destroy.argtypes = [p]
    lib.kemr_idmap_lookup.argtypes = [p, p, p, i64, i32, p]
    lib.kemr_store_write.argtypes = [C.c_char_p, p, i64, i32]
    lib.kemr_store_info.argtypes = [C.c_char_p, C.POINTER(i64), C.POINTER(i32)]
    lib.kemr_store_load.argtypes = [C.c_char_p, i64, i64, p, p]
    lib.kemr_debug_select_stamps.argtypes = [p]
    lib.kemr_debug_mma_plan.argtypes = [i32, i64, i32, i32, i32, i32, i32, i32, p]
    lib.kemr_rank_count.argtypes = [p, i32, p, p, i64, i32, f64, f64, f64, p, p, p, p, p, f64, i64,
                                    p, p, p, sz, i32, p]
    lib.kemr_score_matrix.argtypes = [p, i32, p, p, i64, i32, f32, f32, p, i64, p, sz, i32, p]
    lib.kemr_matrix_rank.argtypes = [p, i32, i64, i64, p, p, p]
    lib.kemr_matrix_topk.argtypes = [p, i32, i64, i64, i32, p, p, p]
    lib.kemr_matrix_fuse.argtypes = [p, p, i32, i64, i64, i32, f32, p, p, p, p]
    lib.kemr_matrix_mlp2.argtypes = [p, p, p, i32, i64, p, p, p, f32, i32, p]
    lib.kemr_infonce_rows.argtypes = [p, p, i32, i32, f32, p, p]
    lib.kemr_metrics_reduce.argtypes = [p, i32, p, i32, p, p, p]
    lib.kemr_metrics_reduce_host.argtypes = [p, i32, p, i32, p, p]
    lib.kemr_merge_topk.argtypes = [p, p, i32, i32, i32, p, p, p]
    lib.kemr_merge_topk_strided.argtypes = [p, p, i64, i32, i32, i32, p, p, p]
    lib.kemr_index_create.argtypes = [p, p, i64, i32, i32, i32, C.POINTER(p)]
    lib.kemr_index_destroy.argtypes = [p]
    lib.kemr_set_scan_done_event.argtypes = [p]
    lib.kemr_set_phase_stamps.argtypes = [p]
    lib.kemr_peer_create.argtypes = [i32, i32, i32, i32, C.POINTER(p), p]
    lib.kemr_peer_connect.argtypes = [p, p]
    lib.kemr_peer_connect_pointers.argtypes = [p, p]
    lib.kemr_peer_local_buffer.argtypes = [p]
    lib.kemr_peer_local_buffer.restype = p
    lib.kemr_peer_destroy.argtypes = [p]
    lib.kemr_peer_begin.argtypes = [p, p]
    lib.kemr_peer_merge.argtypes = [p, i32, i32, p, p, p]
    lib.kemr_peer_gather.argtypes = [p, i32, i32, p, p, p]
    lib.kemr_index_search_host.argtypes = [p, p, i32, i32, f64, f64, f64, p, p, p, i32, p, p, p]
    lib.kemr_index_submit_host.argtypes = [p, p, i32, i32, f64, f64, f64, p, p, p, i32, p, p, p]
    lib.kemr_index_wait.argtypes = [p]
    lib.kemr_index_share.argtypes = [p, C.POINTER(C.c_void_p)]
    lib.kemr_index_search_host_bf16.argtypes = [p, p, i32, f64, f64, f64, p, p, p, i32, p, p, p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("kemr_last_error", "kemr_workspace_bytes", "kemr_abi_version", "kemr_hits_workspace_bytes",
                        "kemr_peer_local_buffer"):
            fn.restype = i32


def load():
    """Load libkemr.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KemrError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C knowledge_enhanced_multimodal_retrieval_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        missing = [n for n in EXPORTS if not hasattr(lib, n)]
        if missing:
            raise KemrError(f"libkemr.so lacks symbols declared in include/kemr.h: {missing}")
        _declare(lib)
        if lib.kemr_abi_version() != ABI_VERSION:
            raise KemrError("libkemr.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc: int):
    if rc != KEMR_OK:
        msg = load().kemr_last_error().decode("utf-8", "replace")
        raise KemrError(f"libkemr error {rc}: {msg}")
