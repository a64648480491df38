"""CPU checks of the per-record logic shared with the CUDA kernels (ptx_core.cuh), via
tests/hostcheck.cpp: the same inline code the kernels run, instantiated with a plain-array
sink, compared with the C++ oracle.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import LABEL_U, NASTY, NASTY_DUP, assert_cpu_matches_py, dataset_graphs, run_cpu_oracle, synth
from test_oracle import kat_inputs, load_kats

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostcheck.so")


def lib():
    src = os.path.join(HERE, "hostcheck.cpp")
    core = os.path.join(HERE, "..", "pantax_b200", "csrc", "ptx_core.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(core), os.path.getmtime(core.replace('ptx_core', 'ptx_fast'))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, src])
    return C.CDLL(SO)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def run_hostcheck(ranges, graphs, gaf: bytes, oracle, stage_lim=0, use_stash=1):
    """Feeds hostcheck the graph + the oracle's trio table (trio build is a separate kernel)."""
    L = lib()
    S = len(ranges)
    rstart = np.array([r[1] for r in ranges], dtype=np.int64)
    rend = np.array([r[2] for r in ranges], dtype=np.int64)
    order = np.argsort(rstart, kind="stable").astype(np.uint32)
    disjoint = int(all(rend[order[i]] < rstart[order[i + 1]] for i in range(S - 1)) and all(rend >= rstart))
    node_base = np.full(S, -1, dtype=np.int64)
    lens, keys, tbase = [], [], []
    N = 0
    T = 0
    for s, g in enumerate(graphs):
        if g is None:
            tbase.append(T)
            continue
        node_base[s] = N
        lens.append(np.asarray(g[0], dtype=np.uint32))
        k, _l, _o = oracle.trio_table(s)
        keys.append((k + N).astype(np.uint32))
        tbase.append(T)
        T += len(k)
        N += len(g[0])
    tbase.append(T)
    ln = np.ascontiguousarray(np.concatenate(lens)) if lens else np.zeros(1, np.uint32)
    tk = np.ascontiguousarray(np.concatenate(keys).reshape(-1)) if T else np.zeros(3, np.uint32)
    nrec_cap = gaf.count(b"\n") + 2
    labels = np.zeros(nrec_cap, dtype=np.uint32)
    nrec = C.c_int64(0)
    hist = np.zeros((S, 4), dtype=np.int64)
    bases = np.zeros(max(N, 1), dtype=np.int64)
    cov = np.zeros(max(N, 1), dtype=np.uint64)
    tb = np.zeros(max(T, 1), dtype=np.int64)
    err = np.zeros(S, dtype=np.uint32)
    uniq = C.c_int(0)
    nover = C.c_int64(0)
    nfast = C.c_int64(0)
    buf = (C.c_char * len(gaf)).from_buffer_copy(gaf)
    rc = L.hostcheck_run(buf, C.c_uint64(len(gaf)), S, _p(rstart, C.c_int64), _p(rend, C.c_int64), _p(node_base, C.c_int64),
                    _p(order, C.c_uint32), disjoint, C.c_int64(N), _p(ln, C.c_uint32), C.c_int64(T), _p(tk, C.c_uint32),
                    C.c_uint32(stage_lim), use_stash, _p(labels, C.c_uint32), C.byref(nrec), _p(hist, C.c_int64), _p(bases, C.c_int64),
                    _p(cov, C.c_uint64), _p(tb, C.c_int64), _p(err, C.c_uint32), C.byref(uniq), C.byref(nover), C.byref(nfast))
    assert rc == 0, f"hostcheck_run failed with code {rc} (9 = parse_head/parse_tail disagree with parse_record, 10 = fast_parse disagrees with parse_record, 11 = classify16 bitmaps wrong, 12 = classify_pivots disagrees with classify)"
    return dict(labels=labels[:nrec.value], hist=hist, bases=bases, cov=cov, trio_bases=tb, err=err,
                ids_unique=bool(uniq.value), node_base=node_base, tbase=tbase, n_overflow=nover.value, n_fast=nfast.value)


def assert_hostcheck_matches(ranges, graphs, gaf, stage_lim=0, use_stash=1):
    o = run_cpu_oracle(ranges, graphs, gaf, threads=2)
    if use_stash:  # the re-parse path (no stash) must give the same answer
        assert_hostcheck_matches(ranges, graphs, gaf, stage_lim, use_stash=0)
    h = run_hostcheck(ranges, graphs, gaf, o, stage_lim, use_stash)
    np.testing.assert_array_equal(h["labels"], o.labels())
    np.testing.assert_array_equal(h["hist"], o.species_counts())
    assert h["ids_unique"] == o.ids_unique
    for s, g in enumerate(graphs):
        if g is None:
            continue
        assert bool(h["err"][s] & 1) == bool(o.species_error(s))
        if o.species_error(s):
            continue
        b = h["node_base"][s]
        n = len(g[0])
        np.testing.assert_array_equal(h["bases"][b:b + n], o.node_bases(s))
        np.testing.assert_array_equal(h["cov"][b:b + n], o.node_cov(s))
        np.testing.assert_array_equal(h["trio_bases"][h["tbase"][s]:h["tbase"][s + 1]], o.trio_bases(s))
    return h, o


@pytest.mark.parametrize("case", load_kats(), ids=lambda c: c["name"])
def test_core_on_kats(case):
    ranges, graphs, gaf = kat_inputs(case)
    assert_hostcheck_matches(ranges, graphs, gaf)


@pytest.mark.parametrize("params,seed", [(synth.GafParams(), 1), (NASTY, 2), (NASTY_DUP, 3)])
def test_core_on_synthetic_short_reads(params, seed):
    ds = synth.Dataset(200 + seed, [3000, 800, 1200, 40], [6, 1, 3, 2])
    gaf = ds.gaf(seed, 0, 5000, params)
    assert_hostcheck_matches(ds.ranges(), dataset_graphs(ds), gaf)


def test_core_on_long_reads_with_small_window():
    """A 200-byte parse window makes most HiFi lines overflow, exercising the re-parse path."""
    ds = synth.Dataset(8, [4000, 2500], [4, 2], backbone_mean=300)
    gaf = ds.gaf(6, 0, 300, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.2))
    h, _ = assert_hostcheck_matches(ds.ranges(), dataset_graphs(ds), gaf, stage_lim=200)
    assert h["n_overflow"] > 0


def test_core_handcrafted_dialect_edges():
    ranges = [("a", 1, 50), ("b", 51, 80)]
    graphs = [(np.full(50, 7, dtype=np.int64), [np.arange(50, dtype=np.uint64), np.array([3, 2, 1, 2, 3, 9], dtype=np.uint64)], ["p", "q"]),
              (np.full(30, 5, dtype=np.int64), [np.arange(30, dtype=np.uint64)], ["r"])]
    lines = [
        b"r1\t50\t0\t50\t+\t>3>4>5\t21\t2\t18\t16\t16\t60\ttp:A:P",         # plain
        b"r2\t50\t0\t50\t+\t>4>3>2>3>4\t35\t1\t30\t29\t29\t60",              # non-monotone walk with repeats
        b"r3\t50\t0\t50\t+\t<10<9\t14\t0\t14\t14\t14\t5\r",                  # CRLF line end
        b"r4\t+50\t0\t50\t+\t>60>61\t10\t-1\t4\t4\t4\t60",                   # signed ints; negative start
        b"r5\t5x\t0\t50\t+\t>7\t7\t1\t3\t2\t2\tabc",                         # junk in int columns -> null
        b"r6\t50\t0\t50\t+\t>8>9",                                           # short row: c7.. null
        b"",                                                                 # empty line
        b"\r",                                                               # CR-only line
        b"@HD\tVN:1.0",                                                      # comment
        b"r7\t50\t0\t50\t+\t>0000000000000000000012>13\t14\t0\t9\t9\t9\t60",  # 22-digit run dropped (rcls.rs:244)
        b"r8\t50\t0\t50\t+\t>49>50>51\t19\t0\t19\t19\t19\t60",               # spans two species -> U
        b"r9\t50\t0\t50\t+\tchr1_12\t7\t0\t5\t5\t5\t3",                      # digit runs inside a name: 1 and 12
        b"*\t50\t0\t50\t+\t>20>21>22\t21\t7\t20\t13\t13\t60",               # '*' read id is a literal id
        b"r10\t50\t0\t50\t+\t>30>31\t14\t8\t20\t12\t12\t60",                 # start 8 > len 7 -> profile.rs:854
    ]
    gaf = b"\n".join(lines)  # no trailing newline on purpose
    # the last record trips the reference's assert: species a is in error for both
    h, o = assert_hostcheck_matches(ranges, graphs, gaf)
    assert o.species_error(0) == 1
    gaf_ok = b"\n".join(lines[:-1]) + b"\n"
    h, o = assert_hostcheck_matches(ranges, graphs, gaf_ok)
    assert o.n_records == 10
    assert h["labels"].tolist() == [0, 0, 0, 1, 0, 0, 0, LABEL_U, 0, 0]
    assert_cpu_matches_py(ranges, graphs, gaf_ok)  # and the naive Python restatement agrees on the dialect edges


def test_id_hash_has_no_collisions_on_realistic_ids():
    L = lib()
    L.hostcheck_idhash.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    seen = {}
    lo, hi = C.c_uint64(), C.c_uint32()
    ids = [b"S%dR%d/%d" % (s, r, m) for s in range(3) for r in range(20000) for m in (1, 2)]
    ids += [b"m64011_190830_220126/%d/ccs" % i for i in range(20000)] + [b"", b"a", b"ab", b"abc", b"abcd", b"abcde"]
    lo32 = set()
    for i in ids:
        L.hostcheck_idhash(i, len(i), C.byref(lo), C.byref(hi))
        key = (lo.value, hi.value)
        assert key not in seen, (i, seen.get(key))
        seen[key] = i
        assert lo.value != 0
        lo32.add(lo.value & 0xFFFFFFFF)
    # even 32 bits of it should be nearly collision free on 140k ids (birthday bound ~2.3 expected)
    assert len(ids) - len(lo32) < 20
