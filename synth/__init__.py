"""ctypes wrapper around synth/synth.cpp (test / bench infrastructure).

`Dataset` owns a synthetic pangenome (species ranges + per-species graphs) and
generates GAF byte streams of the shapes BASELINE.json names.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsynth.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-o", _SO, src]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.synth_create.restype = C.c_void_p
        L.synth_create.argtypes = [C.c_uint64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_double, C.c_int]
        L.synth_destroy.argtypes = [C.c_void_p]
        L.synth_total_nodes.restype = C.c_int64
        L.synth_total_nodes.argtypes = [C.c_void_p]
        L.synth_species_start.restype = C.c_int64
        L.synth_species_start.argtypes = [C.c_void_p, C.c_int]
        L.synth_species_nodes.restype = C.c_int64
        L.synth_species_nodes.argtypes = [C.c_void_p, C.c_int]
        L.synth_species_haps.restype = C.c_int
        L.synth_species_haps.argtypes = [C.c_void_p, C.c_int]
        L.synth_species_len.restype = C.POINTER(C.c_int32)
        L.synth_species_len.argtypes = [C.c_void_p, C.c_int]
        L.synth_path_size.restype = C.c_int64
        L.synth_path_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.synth_path_nodes.restype = C.POINTER(C.c_uint32)
        L.synth_path_nodes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.synth_gaf.restype = C.c_void_p
        L.synth_gaf.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.c_int,
                                C.POINTER(C.c_int64)]
        L.synth_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


@dataclass
class GafParams:
    long_reads: bool = False
    read_len: int = 150
    long_mean: float = 15000.0
    long_sigma: float = 0.3
    p_unmapped: float = 0.001
    p_star_c9: float = 0.0005
    p_neg_single: float = 0.0005
    p_chimera: float = 0.005
    p_dup_same: float = 0.0
    p_dup_other: float = 0.0
    p_secondary: float = 0.0
    p_comment: float = 0.0
    id_pair_suffix: bool = True

    def vec(self):
        v = [float(self.long_reads), float(self.read_len), self.long_mean, self.long_sigma, self.p_unmapped,
             self.p_star_c9, self.p_neg_single, self.p_chimera, self.p_dup_same, self.p_dup_other,
             self.p_secondary, self.p_comment, float(self.id_pair_suffix)]
        return (C.c_double * len(v))(*v)


def hap_name(s: int, h: int) -> str:
    return "GCF_%09d.1" % (s * 1000 + h)


class Dataset:
    def __init__(self, seed: int, nodes_per_species: Sequence[int], haps_per_species: Sequence[int], threads: int = 0,
                 backbone_mean: float = 64.0):
        L = lib()
        n = len(nodes_per_species)
        self.n_species = n
        self.threads = threads or (os.cpu_count() or 1)
        a = (C.c_int64 * n)(*[int(x) for x in nodes_per_species])
        b = (C.c_int * n)(*[int(x) for x in haps_per_species])
        self._h = C.c_void_p(L.synth_create(C.c_uint64(seed), n, a, b, float(backbone_mean), self.threads))
        self.seed = seed
        self.taxids = [str(1000 + s) for s in range(n)]

    def __del__(self):
        try:
            if self._h:
                lib().synth_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def total_nodes(self) -> int:
        return lib().synth_total_nodes(self._h)

    def ranges(self) -> List[Tuple[str, int, int]]:
        """(taxid, start, end) 1-based inclusive global node ids, file order."""
        L = lib()
        out = []
        for s in range(self.n_species):
            st = L.synth_species_start(self._h, s)
            out.append((self.taxids[s], st, st + L.synth_species_nodes(self._h, s) - 1))
        return out

    def nodes_len(self, s: int) -> np.ndarray:
        L = lib()
        n = L.synth_species_nodes(self._h, s)
        return np.ctypeslib.as_array(L.synth_species_len(self._h, s), shape=(n,)).astype(np.int64)

    def paths(self, s: int) -> List[Tuple[str, np.ndarray]]:
        """[(hap name, local node ids uint64)] in name order."""
        L = lib()
        out = []
        for h in range(L.synth_species_haps(self._h, s)):
            n = L.synth_path_size(self._h, s, h)
            arr = np.ctypeslib.as_array(L.synth_path_nodes(self._h, s, h), shape=(n,)).astype(np.uint64)
            out.append((hap_name(s, h), arr))
        return out

    def gaf(self, gseed: int, r0: int, r1: int, params: GafParams = GafParams()) -> bytes:
        buf, n = self.gaf_raw(gseed, r0, r1, params)
        try:
            return C.string_at(buf, n)
        finally:
            lib().synth_free(buf)

    def gaf_raw(self, gseed: int, r0: int, r1: int, params: GafParams = GafParams()):
        """Returns (void* malloc'ed buffer, size); caller frees with lib().synth_free."""
        n = C.c_int64(0)
        p = lib().synth_gaf(self._h, C.c_uint64(gseed), r0, r1, params.vec(), self.threads, C.byref(n))
        return C.c_void_p(p), n.value
