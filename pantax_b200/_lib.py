"""ctypes binding of include/pantax_gpu.h (every exported symbol, exact signatures)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PANTAX_GPU_LIB") or os.path.join(HERE, "libpantax_gpu.so")  # (override: A/B measurements of builds)

ERRORS = {
    0: "PTX_OK", -1: "PTX_E_INVALID", -2: "PTX_E_CUDA", -3: "PTX_E_NOMEM", -4: "PTX_E_STATE", -5: "PTX_E_NODE_ORDER",
    -6: "PTX_E_ZERO_LEN", -7: "PTX_E_START_GT_LEN", -8: "PTX_E_NVERT_MISMATCH", -9: "PTX_E_NO_GRAPH", -10: "PTX_E_NCCL",
    -11: "PTX_E_IO", -12: "PTX_E_RANGE", -13: "PTX_E_UNSUPPORTED",
}


class PantaxGpuError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        self.name = ERRORS.get(code, str(code))
        super().__init__(f"{self.name}: {msg}" if msg else self.name)


vp, i64, u64, u32, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
P = C.POINTER

# name -> (restype, argtypes): the full surface of include/pantax_gpu.h
SIGNATURES = {
    "ptx_create": (i32, [i32, P(vp)]),
    "ptx_destroy": (None, [vp]),
    "ptx_last_error": (C.c_char_p, [vp]),
    "ptx_version": (C.c_char_p, []),
    "ptx_set_ranges": (i32, [vp, i32, P(C.c_char_p), P(i64), P(i64)]),
    "ptx_upload_graph": (i32, [vp, i32, P(i64), i64, P(u64), P(u64), i64]),
    "ptx_upload_graph_gfa": (i32, [vp, i32, vp, C.c_size_t]),
    "ptx_species_graph": (i32, [vp, i32, P(i64), P(u64), P(u64)]),
    "ptx_species_path_steps": (i64, [vp, i32]),
    "ptx_species_path_name": (i32, [vp, i32, i64, C.c_char_p, C.c_size_t]),
    "ptx_commit_graphs": (i32, [vp]),
    "ptx_reserve": (i32, [vp, i64]),
    "ptx_host_alloc": (i32, [C.c_size_t, P(vp)]),
    "ptx_host_free": (i32, [vp]),
    "ptx_ingest_gaf": (i32, [vp, vp, C.c_size_t, i32]),
    "ptx_gaf_buffer_alloc": (i32, [vp, C.c_size_t, P(i32), P(vp)]),
    "ptx_ingest_gaf_device": (i32, [vp, i32, C.c_size_t]),
    "ptx_ingest_labels": (i32, [vp, C.POINTER(C.c_uint32), C.c_int64]),
    "ptx_finalize": (i32, [vp]),
    "ptx_reset": (i32, [vp]),
    "ptx_rewind": (i32, [vp]),
    "ptx_num_records": (i64, [vp]),
    "ptx_num_species": (i32, [vp]),
    "ptx_ids_unique": (i32, [vp]),
    "ptx_read_labels": (i32, [vp, P(u32)]),
    "ptx_species_counts": (i32, [vp, P(i64)]),
    "ptx_equal_length": (i32, [vp, P(i32), P(i64)]),
    "ptx_species_nodes": (i64, [vp, i32]),
    "ptx_species_paths": (i64, [vp, i32]),
    "ptx_species_trios": (i64, [vp, i32]),
    "ptx_node_bases": (i32, [vp, i32, P(i64)]),
    "ptx_node_cov": (i32, [vp, i32, P(u64)]),
    "ptx_node_depth": (i32, [vp, i32, P(C.c_double)]),
    "ptx_trio_bases": (i32, [vp, i32, P(i64)]),
    "ptx_trio_depth": (i32, [vp, i32, P(C.c_double)]),
    "ptx_trio_table": (i32, [vp, i32, P(u64), P(i64), P(u32)]),
    "ptx_trio_ref_order": (i32, [P(u64), P(u64), i64, P(u64), i64, P(u64)]),
    "ptx_path_sums": (i32, [vp, i32, P(i64), P(i64)]),
    "ptx_hap_trio_counts": (i32, [vp, i32, P(i64), P(i64)]),
    "ptx_filter_gaf": (i32, [vp, vp, C.c_size_t, P(u64), i64, P(i64)]),
    "ptx_comm_unique_id": (i32, [vp]),
    "ptx_comm_init": (i32, [vp, i32, i32, vp]),
    "ptx_create_multi": (i32, [P(i32), i32, i64, P(vp)]),
    "ptx_finalize_multi": (i32, [P(vp), i32]),
    "ptx_stats_json": (i32, [vp, C.c_char_p, C.c_size_t]),
    "ptx_timing": (i32, [vp, P(C.c_double), P(C.c_double), P(i64)]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """Loads libpantax_gpu.so and types every symbol.  Raises if the library is missing:
    build it with `python -m pantax_b200.build` (nvcc, sm_100a).  No fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise PantaxGpuError(-2, f"{path} not built; run `python -m pantax_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
