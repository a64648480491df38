"""Shared helpers for the parity tests (oracles are the checkers, never the product)."""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import synth  # noqa: E402
from oracle import cpu as ocpu  # noqa: E402
from oracle import pantax_oracle as opy  # noqa: E402

LABEL_U = 0xFFFFFFFF


def dataset_graphs(ds: "synth.Dataset"):
    """[(nodes_len int64[n], [path uint64[]...], [hap names])] per species."""
    out = []
    for s in range(ds.n_species):
        ps = ds.paths(s)
        out.append((ds.nodes_len(s), [p for _n, p in ps], [n for n, _p in ps]))
    return out


def py_graph(nodes_len, paths, names) -> "opy.Graph":
    g = opy.Graph([int(x) for x in nodes_len])
    g.paths = OrderedDict((n, [int(v) for v in p]) for n, p in zip(names, paths))
    return g


def run_cpu_oracle(ranges, graphs, gaf: bytes, threads: int = 0, labels=None) -> "ocpu.CpuOracle":
    """`labels`: per-row species indices replacing the classifier (strain-only resume)."""
    o = ocpu.CpuOracle(threads)
    o.set_ranges(ranges)
    for s, g in enumerate(graphs):
        if g is not None:
            o.set_graph(s, g[0], g[1])
    o.prepare_graphs()
    if labels is not None:
        o.set_labels(labels)
    o.run(gaf)
    return o


def run_py_oracle(ranges, graphs, gaf: bytes, labels=None):
    gd = {ranges[s][0]: py_graph(*g) for s, g in enumerate(graphs) if g is not None}
    col = None if labels is None else [opy.UNCLASSIFIED if int(l) == LABEL_U else ranges[int(l)][0] for l in labels]
    return opy.coverage_all_species(gaf, ranges, gd, col)


def assert_cpu_matches_py(ranges, graphs, gaf: bytes, labels=None):
    rows, counts, per = run_py_oracle(ranges, graphs, gaf, labels)
    o = run_cpu_oracle(ranges, graphs, gaf, threads=3, labels=labels)
    if labels is not None:
        assert o.label_out_of_range == rows.label_out_of_range
    name_to_idx = {r[0]: i for i, r in enumerate(ranges)}
    lab = o.labels()
    assert len(rows) == o.n_records
    exp = np.array([name_to_idx.get(r.species, LABEL_U) for r in rows], dtype=np.uint32)
    np.testing.assert_array_equal(lab, exp)
    cnt = o.species_counts()
    for sp, c in counts.items():
        np.testing.assert_array_equal(cnt[name_to_idx[sp]], np.array(c, dtype=np.int64))
    assert sum(int(x) for x in cnt.ravel()) == sum(sum(c) for c in counts.values())  # Python ints: the total over species may pass 2^63 in the fuzzer
    for s, g in enumerate(graphs):
        if g is None:
            continue
        p = per[ranges[s][0]]
        if p.get("error"):
            assert o.species_error(s) != 0
            continue
        assert o.species_error(s) == 0
        keys, tlen, owner = o.trio_table(s)
        np.testing.assert_array_equal(keys, np.array(p["trio_keys"], dtype=np.uint64).reshape(-1, 3))
        np.testing.assert_array_equal(tlen, np.array(p["trio_len"], dtype=np.int64))
        np.testing.assert_array_equal(owner, np.array(p["owner"], dtype=np.uint32))
        np.testing.assert_array_equal(o.node_bases(s), np.array(p["bases"], dtype=np.int64))
        np.testing.assert_array_equal(o.node_cov(s), np.array(p["cov"], dtype=np.uint64))
        np.testing.assert_array_equal(o.trio_bases(s), np.array(p["trio_bases"], dtype=np.int64))
        sc, sl = o.path_sums(s)
        np.testing.assert_array_equal(sc, np.array(p["sum_cov"], dtype=np.int64))
        np.testing.assert_array_equal(sl, np.array(p["sum_len"], dtype=np.int64))
        U, nz = o.hap_trio_counts(s)
        np.testing.assert_array_equal(U, np.array(p["U"], dtype=np.int64))
        np.testing.assert_array_equal(nz, np.array(p["nz"], dtype=np.int64))
    return o, (rows, counts, per)


NASTY = synth.GafParams(p_unmapped=0.02, p_star_c9=0.02, p_neg_single=0.02, p_chimera=0.03, p_comment=0.01)
NASTY_DUP = synth.GafParams(p_unmapped=0.02, p_star_c9=0.02, p_neg_single=0.02, p_chimera=0.03, p_comment=0.01,
                            p_dup_same=0.03, p_dup_other=0.03)
