"""ctypes wrapper for oracle/oracle_cpu.cpp (TEST INFRASTRUCTURE ONLY).

Mirrors the getter surface of include/pantax_gpu.h so tests can compare arrays 1:1.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_cpu.so")
LABEL_U = 0xFFFFFFFF
NULL_I64 = -(2 ** 63)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_cpu.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_cpu.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64p, u64p, u32p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_threads.argtypes = [vp]
        L.orc_set_ranges.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), i64p, i64p]
        L.orc_set_graph.argtypes = [vp, C.c_int, i64p, C.c_int64, u64p, u64p, C.c_int64]
        L.orc_prepare_graphs.restype = C.c_double
        L.orc_prepare_graphs.argtypes = [vp]
        L.orc_run.argtypes = [vp, C.c_void_p, C.c_size_t]
        L.orc_set_labels.argtypes = [vp, u32p, C.c_int64]
        L.orc_label_out_of_range.argtypes = [vp]
        L.orc_n_records.restype = C.c_int64
        L.orc_n_records.argtypes = [vp]
        L.orc_ids_unique.argtypes = [vp]
        L.orc_mixed_dropped.restype = C.c_int64
        L.orc_mixed_dropped.argtypes = [vp]
        L.orc_times.argtypes = [vp, C.POINTER(C.c_double)]
        L.orc_workload_stats.argtypes = [vp, i64p]
        L.orc_labels.argtypes = [vp, u32p]
        L.orc_record_fields.argtypes = [vp, i64p, i64p]
        L.orc_species_counts.argtypes = [vp, i64p]
        L.orc_species_error.argtypes = [vp, C.c_int]
        L.orc_trio_count.restype = C.c_int64
        L.orc_trio_count.argtypes = [vp, C.c_int]
        L.orc_trio_table.argtypes = [vp, C.c_int, u64p, i64p, u32p]
        L.orc_node_bases.argtypes = [vp, C.c_int, i64p]
        L.orc_node_cov.argtypes = [vp, C.c_int, u64p]
        L.orc_trio_bases.argtypes = [vp, C.c_int, i64p]
        L.orc_path_sums.argtypes = [vp, C.c_int, i64p, i64p]
        L.orc_hap_trio_counts.argtypes = [vp, C.c_int, i64p, i64p]
        L.orc_filter_gaf.restype = C.c_int64
        L.orc_filter_gaf.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_int64]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class CpuOracle:
    def __init__(self, threads: int = 0):
        self._L = lib()
        self._h = C.c_void_p(self._L.orc_create(threads))
        self.n_nodes: List[int] = []
        self.n_haps: List[int] = []
        self._keep = []

    def __del__(self):
        try:
            if self._h:
                self._L.orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def threads(self) -> int:
        return self._L.orc_threads(self._h)

    def set_ranges(self, ranges: Sequence[Tuple[str, int, int]]):
        S = len(ranges)
        names = (C.c_char_p * S)(*[r[0].encode() for r in ranges])
        st = np.array([r[1] for r in ranges], dtype=np.int64)
        en = np.array([r[2] for r in ranges], dtype=np.int64)
        self._L.orc_set_ranges(self._h, S, names, _p(st, C.c_int64), _p(en, C.c_int64))
        self.S = S
        self.n_nodes = [0] * S
        self.n_haps = [0] * S

    def set_graph(self, s: int, nodes_len: np.ndarray, paths: Sequence[np.ndarray]):
        nl = np.ascontiguousarray(nodes_len, dtype=np.int64)
        off = np.zeros(len(paths) + 1, dtype=np.uint64)
        for i, p in enumerate(paths):
            off[i + 1] = off[i] + len(p)
        flat = np.concatenate([np.asarray(p, dtype=np.uint64) for p in paths]) if paths else np.zeros(0, np.uint64)
        flat = np.ascontiguousarray(flat, dtype=np.uint64)
        rc = self._L.orc_set_graph(self._h, s, _p(nl, C.c_int64), len(nl), _p(off, C.c_uint64), _p(flat, C.c_uint64),
                                   len(paths))
        if rc != 0:
            raise ValueError("nvert mismatch")
        self.n_nodes[s] = len(nl)
        self.n_haps[s] = len(paths)

    def prepare_graphs(self) -> float:
        return self._L.orc_prepare_graphs(self._h)

    def set_labels(self, labels):
        """Strain-only resume: per-row species labels replace the classifier's (profile.rs:3367-3385)."""
        a = np.ascontiguousarray(labels, dtype=np.uint32)
        self._L.orc_set_labels(self._h, _p(a, C.c_uint32), a.size)

    @property
    def label_out_of_range(self) -> bool:
        return bool(self._L.orc_label_out_of_range(self._h))

    def run(self, data, size=None):
        """`data`: bytes or an integer address (+ size)."""
        if isinstance(data, (bytes, bytearray)):
            self._keep = data  # Rec holds pointers into it
            buf = (C.c_char * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else None
            self._keep = buf
            self._L.orc_run(self._h, C.addressof(buf), len(data))
        else:
            self._L.orc_run(self._h, C.c_void_p(data), size)

    def times(self):
        t = (C.c_double * 4)()
        self._L.orc_times(self._h, t)
        return list(t)

    @property
    def n_records(self) -> int:
        return self._L.orc_n_records(self._h)

    @property
    def ids_unique(self) -> bool:
        return bool(self._L.orc_ids_unique(self._h))

    @property
    def mixed_dropped(self) -> int:
        return self._L.orc_mixed_dropped(self._h)

    def workload_stats(self):
        a = np.zeros(4, dtype=np.int64)
        self._L.orc_workload_stats(self._h, _p(a, C.c_int64))
        return dict(cover_reads=int(a[0]), walk_nodes=int(a[1]), trio_windows=int(a[2]), trio_hits=int(a[3]))

    def labels(self) -> np.ndarray:
        out = np.empty(self.n_records, dtype=np.uint32)
        self._L.orc_labels(self._h, _p(out, C.c_uint32))
        return out

    def record_fields(self):
        a = np.empty(self.n_records, dtype=np.int64)
        b = np.empty(self.n_records, dtype=np.int64)
        self._L.orc_record_fields(self._h, _p(a, C.c_int64), _p(b, C.c_int64))
        return a, b

    def species_counts(self) -> np.ndarray:
        out = np.empty((self.S, 4), dtype=np.int64)
        self._L.orc_species_counts(self._h, _p(out, C.c_int64))
        return out

    def species_error(self, s: int) -> int:
        return self._L.orc_species_error(self._h, s)

    def trio_table(self, s: int):
        T = self._L.orc_trio_count(self._h, s)
        keys = np.empty((T, 3), dtype=np.uint64)
        ln = np.empty(T, dtype=np.int64)
        ow = np.empty(T, dtype=np.uint32)
        self._L.orc_trio_table(self._h, s, _p(keys, C.c_uint64), _p(ln, C.c_int64), _p(ow, C.c_uint32))
        return keys, ln, ow

    def node_bases(self, s: int) -> np.ndarray:
        out = np.empty(self.n_nodes[s], dtype=np.int64)
        self._L.orc_node_bases(self._h, s, _p(out, C.c_int64))
        return out

    def node_cov(self, s: int) -> np.ndarray:
        out = np.empty(self.n_nodes[s], dtype=np.uint64)
        self._L.orc_node_cov(self._h, s, _p(out, C.c_uint64))
        return out

    def trio_bases(self, s: int) -> np.ndarray:
        out = np.empty(self._L.orc_trio_count(self._h, s), dtype=np.int64)
        self._L.orc_trio_bases(self._h, s, _p(out, C.c_int64))
        return out

    def path_sums(self, s: int):
        a = np.empty(self.n_haps[s], dtype=np.int64)
        b = np.empty(self.n_haps[s], dtype=np.int64)
        self._L.orc_path_sums(self._h, s, _p(a, C.c_int64), _p(b, C.c_int64))
        return a, b

    def hap_trio_counts(self, s: int):
        a = np.empty(self.n_haps[s], dtype=np.int64)
        b = np.empty(self.n_haps[s], dtype=np.int64)
        self._L.orc_hap_trio_counts(self._h, s, _p(a, C.c_int64), _p(b, C.c_int64))
        return a, b


def filter_gaf(data: bytes) -> List[bytes]:
    L = lib()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    cap = data.count(b"\n") + 1
    off = np.empty(cap, dtype=np.uint64)
    k = L.orc_filter_gaf(C.addressof(buf), len(data), _p(off, C.c_uint64), cap)
    out = []
    for o in off[:k]:
        o = int(o)
        e = data.find(b"\n", o)
        line = data[o:e if e >= 0 else len(data)]
        out.append(line[:-1] if line.endswith(b"\r") else line)
    return out
