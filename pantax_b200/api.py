"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Names follow the reference (paths relative to /root/reference/pantax/src):
  rcls_profile            rcls.rs:452            GAF -> species label per read
  species_counts          profile.rs:208-297     integer part of species_profiling
  trio_nodes_info         profile.rs:658-740     unique trio table of a species
  get_node_abundances     profile.rs:743-1026    (node_abundance_vec, trio_node_abundance_vec, node_base_cov)
  path_cov_ratio          profile.rs:2705-2729   covered fraction per path
  hap_trio_counts         profile.rs:1112-1135   U_h, nz_h
  trio_ref_order          profile.rs:659-716     the reference's numbering (FxHashSet order) of the unique trio table

Everything numeric is computed by libpantax_gpu.so on the GPU; this file only moves
arrays across ctypes.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import PantaxGpuError, load_library

LABEL_U = 0xFFFFFFFF


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def bind_host_thread_to_gpu(device: int) -> Optional[List[int]]:
    """Restrict the calling process to the CPU cores next to `device` (NVML cpu affinity), so that the pinned GAF buffer
    allocated afterwards lives on the GPU's own NUMA node: with one process per GPU on a two-socket box, H2D copies from
    the far socket share the inter-socket link (round 1: end-to-end efficiency 0.52 at 4 GPUs).  Returns the core list,
    or None when NVML / sched_setaffinity is unavailable (nothing is changed then)."""
    import os

    try:
        import pynvml as n

        n.nvmlInit()
        phys = device
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            phys = int(vis.split(",")[device])
        h = n.nvmlDeviceGetHandleByIndex(phys)
        words = n.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


class PinnedBuffer:
    """Pinned host memory from ptx_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self._L = load_library()
        self._ptr = C.c_void_p()
        rc = self._L.ptx_host_alloc(nbytes, C.byref(self._ptr))
        if rc:
            raise PantaxGpuError(rc, "ptx_host_alloc")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_uint8)), shape=(max(nbytes, 1),))[:nbytes]

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def view(self, dtype, count: int, byte_offset: int = 0) -> np.ndarray:
        """A typed numpy view of part of the buffer (e.g. an `out=` array for the result getters: a pinned
        destination lets the device->host copy run at PCIe speed instead of through pageable staging)."""
        dt = np.dtype(dtype)
        assert byte_offset % dt.itemsize == 0 and byte_offset + count * dt.itemsize <= self.nbytes
        return self.array[byte_offset: byte_offset + count * dt.itemsize].view(dt)

    def free(self):
        if self._ptr:
            self._L.ptx_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PantaxGpu:
    """One context per process per GPU (include/pantax_gpu.h)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        self._h = C.c_void_p()
        rc = self._L.ptx_create(device, C.byref(self._h))
        if rc:
            raise PantaxGpuError(rc, "ptx_create: no usable CUDA device (this library has no CPU fallback)")
        self.taxids: List[str] = []

    @classmethod
    def create_multi(cls, devices: Sequence[int], expected_records_per_device: int = 0) -> List["PantaxGpu"]:
        """ptx_create_multi: one context per device, all in THIS process, already one communicator group
        (rank = position in `devices`).  Finalize them together with `PantaxGpu.finalize_multi`."""
        L = load_library()
        n = len(devices)
        devs = (C.c_int * n)(*[int(d) for d in devices])
        out = (C.c_void_p * n)()
        rc = L.ptx_create_multi(devs, n, int(expected_records_per_device), out)
        if rc:
            raise PantaxGpuError(rc, "ptx_create_multi")
        ctxs = []
        for i in range(n):
            c = cls.__new__(cls)
            c._L = L
            c._h = C.c_void_p(out[i])
            c.taxids = []
            ctxs.append(c)
        return ctxs

    @staticmethod
    def finalize_multi(ctxs: Sequence["PantaxGpu"]):
        L = load_library()
        n = len(ctxs)
        arr = (C.c_void_p * n)(*[c._h for c in ctxs])
        rc = L.ptx_finalize_multi(arr, n)
        if rc:
            msgs = "; ".join((L.ptx_last_error(c._h) or b"").decode() for c in ctxs)
            raise PantaxGpuError(rc, msgs)

    # -- plumbing ------------------------------------------------------------------
    def _ck(self, rc: int):
        if rc:
            raise PantaxGpuError(rc, (self._L.ptx_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._L.ptx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs --------------------------------------------------------------------
    def set_ranges(self, ranges: Sequence[Tuple[str, int, int]]):
        """species_range.txt rows in file order (rcls.rs:40-71)."""
        S = len(ranges)
        names = (C.c_char_p * S)(*[str(r[0]).encode() for r in ranges])
        st = np.array([r[1] for r in ranges], dtype=np.int64)
        en = np.array([r[2] for r in ranges], dtype=np.int64)
        self._ck(self._L.ptx_set_ranges(self._h, S, names, _p(st, C.c_int64), _p(en, C.c_int64)))
        self.taxids = [str(r[0]) for r in ranges]

    def upload_graph(self, species: int, nodes_len: np.ndarray, paths: Sequence[np.ndarray]):
        """types.rs:51-55 Graph; `paths` in hap-name order, local node ids."""
        nl = np.ascontiguousarray(nodes_len, dtype=np.int64)
        off = np.zeros(len(paths) + 1, dtype=np.uint64)
        for i, p in enumerate(paths):
            off[i + 1] = off[i] + len(p)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint64) for p in paths]) if len(paths) else
                                    np.zeros(0, np.uint64), dtype=np.uint64)
        if flat.size == 0:
            flat = np.zeros(1, np.uint64)
        self._ck(self._L.ptx_upload_graph(self._h, species, _p(nl, C.c_int64), len(nl), _p(off, C.c_uint64),
                                          _p(flat, C.c_uint64), len(paths)))

    def upload_graph_gfa(self, species: int, gfa: bytes):
        """profile.rs:466-545 read_gfa(previous = 0), parsed on the device: the species' GFA text instead of arrays."""
        buf = (C.c_uint8 * len(gfa)).from_buffer_copy(gfa)
        self._ck(self._L.ptx_upload_graph_gfa(self._h, species, C.cast(buf, C.c_void_p), len(gfa)))

    def species_graph(self, species: int):
        """(nodes_len int64[n], [path uint64[]...], [haplotype ids]) as the library holds them (what read_gfa returns)."""
        n, H, P = self.n_nodes(species), self.n_paths(species), int(self._L.ptx_species_path_steps(self._h, species))
        nl, off, flat = np.zeros(max(n, 1), np.int64), np.zeros(H + 1, np.uint64), np.zeros(max(P, 1), np.uint64)
        self._ck(self._L.ptx_species_graph(self._h, species, _p(nl, C.c_int64), _p(off, C.c_uint64), _p(flat, C.c_uint64)))
        names = []
        for h in range(H):
            buf = C.create_string_buffer(4096)
            k = self._L.ptx_species_path_name(self._h, species, h, buf, 4096)
            if k < 0:
                self._ck(k)
            names.append(buf.value.decode())
        return nl[:n], [flat[int(off[h]):int(off[h + 1])] for h in range(H)], names

    def commit_graphs(self):
        self._ck(self._L.ptx_commit_graphs(self._h))

    def reserve(self, n_records: int):
        self._ck(self._L.ptx_reserve(self._h, n_records))

    def ingest_gaf(self, data, size: Optional[int] = None, is_last: bool = True):
        """`data`: bytes / numpy uint8 array / PinnedBuffer / integer host address (+size)."""
        if isinstance(data, (bytes, bytearray)):
            buf = np.frombuffer(data, dtype=np.uint8)
            self._ck(self._L.ptx_ingest_gaf(self._h, buf.ctypes.data, buf.size, int(is_last)))
        elif isinstance(data, PinnedBuffer):
            self._ck(self._L.ptx_ingest_gaf(self._h, data.ptr, data.nbytes if size is None else size, int(is_last)))
        elif isinstance(data, np.ndarray):
            a = np.ascontiguousarray(data, dtype=np.uint8)
            self._ck(self._L.ptx_ingest_gaf(self._h, a.ctypes.data, a.size if size is None else size, int(is_last)))
        else:
            self._ck(self._L.ptx_ingest_gaf(self._h, C.c_void_p(int(data)), size, int(is_last)))

    def ingest_labels(self, labels):
        """profile.rs:3367-3385 (strain-only resume): species label per GAF row (index into the ranges, LABEL_U = "U")
        taken from reads_classification.tsv instead of the walk; call before ingest_gaf."""
        a = np.ascontiguousarray(labels, dtype=np.uint32)
        self._ck(self._L.ptx_ingest_labels(self._h, _p(a, C.c_uint32), a.size))

    def gaf_buffer_alloc(self, capacity: int) -> Tuple[int, int]:
        """Returns (buffer_id, device pointer) of a padded device buffer for HBM-resident GAF text."""
        bid = C.c_int()
        ptr = C.c_void_p()
        self._ck(self._L.ptx_gaf_buffer_alloc(self._h, capacity, C.byref(bid), C.byref(ptr)))
        return bid.value, ptr.value

    def ingest_gaf_device(self, buffer_id: int, n: int):
        self._ck(self._L.ptx_ingest_gaf_device(self._h, buffer_id, n))

    def finalize(self):
        self._ck(self._L.ptx_finalize(self._h))

    def reset(self):
        self._ck(self._L.ptx_reset(self._h))

    def rewind(self):
        self._ck(self._L.ptx_rewind(self._h))

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self._L.ptx_comm_init(self._h, n_ranks, rank, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.ptx_comm_unique_id(buf)
        if rc:
            raise PantaxGpuError(rc, "ptx_comm_unique_id")
        return buf.raw

    # -- outputs -------------------------------------------------------------------
    @property
    def num_records(self) -> int:
        return self._L.ptx_num_records(self._h)

    @property
    def ids_unique(self) -> bool:
        return bool(self._L.ptx_ids_unique(self._h))

    def read_labels(self) -> np.ndarray:
        out = np.empty(max(self.num_records, 1), dtype=np.uint32)
        self._ck(self._L.ptx_read_labels(self._h, _p(out, C.c_uint32)))
        return out[: self.num_records]

    def species_counts(self) -> np.ndarray:
        out = np.zeros((len(self.taxids), 4), dtype=np.int64)
        self._ck(self._L.ptx_species_counts(self._h, _p(out, C.c_int64)))
        return out

    def equal_length(self) -> Tuple[bool, int]:
        eq = C.c_int()
        rl = C.c_int64()
        self._ck(self._L.ptx_equal_length(self._h, C.byref(eq), C.byref(rl)))
        return bool(eq.value), rl.value

    def n_nodes(self, s: int) -> int:
        return self._L.ptx_species_nodes(self._h, s)

    def n_paths(self, s: int) -> int:
        return self._L.ptx_species_paths(self._h, s)

    def n_trios(self, s: int) -> int:
        return self._L.ptx_species_trios(self._h, s)

    def _sized(self, n: int, dtype, out=None):
        if out is not None:  # caller-owned destination (may be pinned memory)
            if out.dtype != np.dtype(dtype) or out.size < max(n, 1) or not out.flags.c_contiguous:
                raise ValueError(f"out must be a contiguous {np.dtype(dtype)} array of >= {max(n, 1)} elements")
            return out
        return np.zeros(max(n, 1), dtype=dtype)

    def node_bases(self, s: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.n_nodes(s)
        out = self._sized(n, np.int64, out)
        self._ck(self._L.ptx_node_bases(self._h, s, _p(out, C.c_int64)))
        return out[:n]

    def node_cov(self, s: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.n_nodes(s)
        out = self._sized(n, np.uint64, out)
        self._ck(self._L.ptx_node_cov(self._h, s, _p(out, C.c_uint64)))
        return out[:n]

    def node_depth(self, s: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.n_nodes(s)
        out = self._sized(n, np.float64, out)
        self._ck(self._L.ptx_node_depth(self._h, s, _p(out, C.c_double)))
        return out[:n]

    def trio_bases(self, s: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.n_trios(s)
        out = self._sized(n, np.int64, out)
        self._ck(self._L.ptx_trio_bases(self._h, s, _p(out, C.c_int64)))
        return out[:n]

    def trio_depth(self, s: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.n_trios(s)
        out = self._sized(n, np.float64, out)
        self._ck(self._L.ptx_trio_depth(self._h, s, _p(out, C.c_double)))
        return out[:n]

    def trio_table(self, s: int):
        n = self.n_trios(s)
        keys = np.zeros((max(n, 1), 3), dtype=np.uint64)
        ln = self._sized(n, np.int64)
        ow = self._sized(n, np.uint32)
        self._ck(self._L.ptx_trio_table(self._h, s, _p(keys, C.c_uint64), _p(ln, C.c_int64), _p(ow, C.c_uint32)))
        return keys[:n], ln[:n], ow[:n]

    def path_sums(self, s: int):
        n = self.n_paths(s)
        a, b = self._sized(n, np.int64), self._sized(n, np.int64)
        self._ck(self._L.ptx_path_sums(self._h, s, _p(a, C.c_int64), _p(b, C.c_int64)))
        return a[:n], b[:n]

    def hap_trio_counts(self, s: int):
        n = self.n_paths(s)
        a, b = self._sized(n, np.int64), self._sized(n, np.int64)
        self._ck(self._L.ptx_hap_trio_counts(self._h, s, _p(a, C.c_int64), _p(b, C.c_int64)))
        return a[:n], b[:n]

    def filter_gaf(self, data: bytes):
        """gaf_filter.rs:44-97: returns the kept lines (file order, first qualifying line per read id)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        cap = data.count(b"\n") + 1
        off = np.zeros(cap, dtype=np.uint64)
        n_out = C.c_int64()
        self._ck(self._L.ptx_filter_gaf(self._h, buf.ctypes.data, buf.size, _p(off, C.c_uint64), cap, C.byref(n_out)))
        out = []
        for o in off[: n_out.value]:
            o = int(o)
            e = data.find(b"\n", o)
            line = data[o:e if e >= 0 else len(data)]
            out.append(line[:-1] if line.endswith(b"\r") else line)
        return out

    def timing(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self._L.ptx_timing(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def stats(self) -> dict:
        buf = C.create_string_buffer(2048)
        self._ck(self._L.ptx_stats_json(self._h, buf, 2048))
        return json.loads(buf.value.decode())


# ---- reference-named entry points ------------------------------------------------------
def rcls_profile(ctx: PantaxGpu, gaf, ranges: Optional[Sequence[Tuple[str, int, int]]] = None) -> np.ndarray:
    """rcls.rs:452-458: returns the species label of every GAF row (index into ranges, LABEL_U = "U")."""
    if ranges is not None:
        ctx.set_ranges(ranges)
    ctx.ingest_gaf(gaf, is_last=True)
    ctx.finalize()
    return ctx.read_labels()


def species_counts(ctx: PantaxGpu) -> np.ndarray:
    """profile.rs:219-232/264-277: [read_count, sum(read_len), less_multi, uniq_count] per species."""
    return ctx.species_counts()


def trio_nodes_info(ctx: PantaxGpu, species: int):
    """profile.rs:658-740: (unique trio keys[T,3] canonical local ids, unique_lengths[T], owner hap[T])."""
    return ctx.trio_table(species)


def trio_ref_order(paths: Sequence[np.ndarray], keys3: np.ndarray) -> np.ndarray:
    """order[i] = row of `keys3` (ptx_trio_table) that the reference numbers i: the iteration order of the FxHashSet of
    profile.rs:659-685 restricted to the unique trios (:705-716).  Host-only (ptx_trio_ref_order); `paths` = local node ids per
    hap in name order.  The one place the order shows is the f64 summation order of frequencies_mean (profile.rs:1037-1146)."""
    L = load_library()
    off = np.zeros(len(paths) + 1, dtype=np.uint64)
    for h, p in enumerate(paths):
        off[h + 1] = off[h] + len(p)
    nodes = np.concatenate([np.asarray(p, dtype=np.uint64) for p in paths]) if len(paths) and off[-1] else np.zeros(1, dtype=np.uint64)
    nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
    k = np.ascontiguousarray(np.asarray(keys3, dtype=np.uint64).reshape(-1, 3))
    T = k.shape[0]
    order = np.zeros(max(T, 1), dtype=np.uint64)
    kk = k if T else np.zeros((1, 3), dtype=np.uint64)
    rc = L.ptx_trio_ref_order(_p(off, C.c_uint64), _p(nodes, C.c_uint64), len(paths), _p(kk, C.c_uint64), T, _p(order, C.c_uint64))
    if rc != 0:
        raise PantaxGpuError(rc, "ptx_trio_ref_order: the paths and the trio table do not describe the same unique trios")
    return order[:T].astype(np.int64)


def get_node_abundances(ctx: PantaxGpu, species: int):
    """profile.rs:743-1026: (node_abundance_vec f64[n], trio_node_abundance_vec f64[T], node_base_cov u64[n])."""
    return ctx.node_depth(species), ctx.trio_depth(species), ctx.node_cov(species)


def path_cov_ratio(ctx: PantaxGpu, species: int, paths: Sequence[np.ndarray], nodes_len: np.ndarray, f32: bool = True) -> np.ndarray:
    """profile.rs:2714-2729: covered fraction of every path over its distinct nodes, with the reference's arithmetic.
    The reference multiplies `RowDVector<f32>(node_base_cov)` by the 0/1 incidence matrix: nalgebra does that as one gemv
    per path, a SEQUENTIAL f32 accumulation over the nodes in index order - not the exact integer once a running sum
    passes 2^24 (a path of 1 M nodes x 30 bp does).  So the quotient is formed here on the host from ptx_node_cov in
    exactly that order (f64 for the CBC variant, profile.rs:1952-1977, where the sums stay exact).  `paths`: local node
    ids per hap in name order (the Graph the caller uploaded), `nodes_len`: its node lengths.  ptx_path_sums still
    returns the exact integer sums."""
    ft = np.float32 if f32 else np.float64
    cov = ctx.node_cov(species).astype(ft)
    ln = np.asarray(nodes_len).astype(ft)
    out = np.empty(len(paths), dtype=np.float64)
    for h, p in enumerate(paths):
        d = np.unique(np.asarray(p, dtype=np.int64))  # distinct nodes, ascending index
        if d.size == 0:
            out[h] = np.nan
            continue
        # np.add.accumulate is a strict left-to-right running sum in the given dtype
        sc = np.add.accumulate(cov[d], dtype=ft)[-1]
        sl = np.add.accumulate(ln[d], dtype=ft)[-1]
        out[h] = float(ft(sc / sl))
    return out


def filter_max_alignment_mt(ctx: PantaxGpu, gaf: bytes):
    """gaf_filter.rs:44-97 (long-read pre-filter): the lines the reference writes to *_filtered.gaf."""
    return ctx.filter_gaf(gaf)


def hap_trio_counts(ctx: PantaxGpu, species: int):
    """profile.rs:1114-1135: (#unique trios owned, # of those with abundance > 0) per hap."""
    return ctx.hap_trio_counts(species)
