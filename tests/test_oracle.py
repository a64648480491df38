"""Pins for the oracles: hand-derived KATs (tests/golden/kat_cases.json) and the
two independent restatements (Python naive vs multithreaded C++) against each other."""
import json
import os

import numpy as np
import pytest

from common import (LABEL_U, NASTY, NASTY_DUP, assert_cpu_matches_py, dataset_graphs, ocpu, opy, run_cpu_oracle,
                    run_py_oracle, synth)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat_cases.json")


def load_kats():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def kat_inputs(case):
    ranges = [tuple(r) for r in case["ranges"]]
    graphs = []
    for name, _s, _e in ranges:
        g = case["graphs"].get(name)
        if g is None:
            graphs.append(None)
            continue
        names = sorted(g["paths"], key=lambda k: k.encode())
        graphs.append((np.array(g["nodes_len"], dtype=np.int64),
                       [np.array(g["paths"][n], dtype=np.uint64) for n in names], names))
    gaf = ("\n".join(case["gaf_lines"]) + "\n").encode()
    return ranges, graphs, gaf


@pytest.mark.parametrize("case", load_kats(), ids=lambda c: c["name"])
def test_python_oracle_matches_hand_derived_kats(case):
    ranges, graphs, gaf = kat_inputs(case)
    rows, counts, per = run_py_oracle(ranges, graphs, gaf)
    exp = case["expect"]
    assert [r.species for r in rows] == exp["labels"]
    assert {k: v for k, v in counts.items()} == exp["counts"]
    for name, _s, _e in ranges:
        if name not in exp:
            continue
        e, p = exp[name], per[name]
        assert [list(k) for k in p["trio_keys"]] == e["trio_keys"]
        assert p["trio_len"] == e["trio_len"]
        assert p["owner"] == e["trio_owner"]
        assert p["bases"] == e["bases"]
        assert p["cov"] == e["cov"]
        assert p["trio_bases"] == e["trio_bases"]
        assert p["U"] == e["hap_U"]
        assert p["nz"] == e["hap_nz"]
        assert p["sum_cov"] == e["path_sum_cov"]
        assert p["sum_len"] == e["path_sum_len"]


@pytest.mark.parametrize("case", load_kats(), ids=lambda c: c["name"])
def test_cpp_oracle_matches_hand_derived_kats(case):
    ranges, graphs, gaf = kat_inputs(case)
    o = run_cpu_oracle(ranges, graphs, gaf, threads=2)
    exp = case["expect"]
    idx = {r[0]: i for i, r in enumerate(ranges)}
    np.testing.assert_array_equal(o.labels(), np.array([idx.get(l, LABEL_U) for l in exp["labels"]], dtype=np.uint32))
    cnt = o.species_counts()
    for k, v in exp["counts"].items():
        assert cnt[idx[k]].tolist() == v
    assert o.ids_unique == exp["ids_unique"]
    for name, _s, _e in ranges:
        if name not in exp:
            continue
        e, s = exp[name], idx[name]
        keys, tlen, owner = o.trio_table(s)
        assert keys.tolist() == e["trio_keys"]
        assert tlen.tolist() == e["trio_len"]
        assert owner.tolist() == e["trio_owner"]
        assert o.node_bases(s).tolist() == e["bases"]
        assert o.node_cov(s).tolist() == e["cov"]
        assert o.trio_bases(s).tolist() == e["trio_bases"]
        U, nz = o.hap_trio_counts(s)
        assert U.tolist() == e["hap_U"] and nz.tolist() == e["hap_nz"]
        sc, sl = o.path_sums(s)
        assert sc.tolist() == e["path_sum_cov"] and sl.tolist() == e["path_sum_len"]


@pytest.mark.parametrize("params,seed", [(synth.GafParams(), 1), (NASTY, 2), (NASTY_DUP, 3)])
def test_cpp_and_python_oracles_agree_short_reads(params, seed):
    ds = synth.Dataset(100 + seed, [3000, 800, 1200, 40], [6, 1, 3, 2])
    gaf = ds.gaf(seed, 0, 3000, params)
    assert_cpu_matches_py(ds.ranges(), dataset_graphs(ds), gaf)


def test_cpp_and_python_oracles_agree_long_reads_and_filter():
    ds = synth.Dataset(7, [4000, 2500], [4, 2], backbone_mean=300)
    gl = ds.gaf(6, 0, 200, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.3))
    o, _ = assert_cpu_matches_py(ds.ranges(), dataset_graphs(ds), gl)
    assert not o.ids_unique
    a = opy.filter_max_alignment(gl)
    b = ocpu.filter_gaf(gl)
    assert a == b and 0 < len(a) < gl.count(b"\n")
    # after the filter every id is unique (gaf_filter.rs:85 "only one line per read_id")
    ids = [l.split(b"\t")[0] for l in a]
    assert len(ids) == len(set(ids))


def _supplied_labels(ranges, graphs, gaf, seed, p_to_u=0.05, p_wrong=0.0):
    """Species column for the strain-only resume: the classifier's labels with some rows set to "U" (a stricter
    binning file) and optionally some rows moved to a wrong species (walk outside that species' range)."""
    lab = run_cpu_oracle(ranges, graphs, gaf).labels().copy()
    rng = np.random.default_rng(seed)
    lab[rng.random(lab.size) < p_to_u] = LABEL_U
    if p_wrong:
        hit = (rng.random(lab.size) < p_wrong) & (lab != LABEL_U)
        lab[hit] = (lab[hit] + 1) % len(ranges)
    return lab


def test_supplied_species_column_python_vs_cpp():
    # profile.rs:3367-3385: --strain without --species takes column 3 of reads_classification.tsv as the species
    ds = synth.Dataset(61, [3000, 800, 1200], [6, 1, 3])
    gaf = ds.gaf(5, 0, 3000, NASTY_DUP)
    graphs = dataset_graphs(ds)
    lab = _supplied_labels(ds.ranges(), graphs, gaf, 1)
    o, _ = assert_cpu_matches_py(ds.ranges(), graphs, gaf, labels=lab)
    assert not o.label_out_of_range
    np.testing.assert_array_equal(o.labels(), lab)
    ref = run_cpu_oracle(ds.ranges(), graphs, gaf)
    assert o.species_counts()[:, 0].sum() < ref.species_counts()[:, 0].sum()
    # rows moved to a species whose range does not hold their walk are reported and treated as "U"
    lab2 = _supplied_labels(ds.ranges(), graphs, gaf, 2, p_wrong=0.02)
    o2, _ = assert_cpu_matches_py(ds.ranges(), graphs, gaf, labels=lab2)
    assert o2.label_out_of_range


def test_oracle_invariant_under_record_permutation_and_thread_count():
    ds = synth.Dataset(9, [2000, 600], [5, 2])
    gaf = ds.gaf(4, 0, 2000, NASTY)
    lines = gaf.split(b"\n")[:-1]
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(lines))
    gaf2 = b"\n".join(lines[i] for i in perm) + b"\n"
    graphs = dataset_graphs(ds)
    a = run_cpu_oracle(ds.ranges(), graphs, gaf, threads=1)
    b = run_cpu_oracle(ds.ranges(), graphs, gaf2, threads=4)
    np.testing.assert_array_equal(a.species_counts(), b.species_counts())
    for s in range(2):
        np.testing.assert_array_equal(a.node_bases(s), b.node_bases(s))
        np.testing.assert_array_equal(a.node_cov(s), b.node_cov(s))
        np.testing.assert_array_equal(a.trio_bases(s), b.trio_bases(s))


def test_species_profiling_tail_matches_readme_shape():
    """profile.rs:299-349 float tail: abundance = bases/len normalised, sorted descending."""
    ds = synth.Dataset(5, [2000, 600, 900], [5, 2, 1])
    gaf = ds.gaf(4, 0, 3000)
    rows = opy.rcls_profile(gaf, ds.ranges())
    table = opy.species_profiling(rows, {"1000": 81000.0, "1001": 40000.0, "1002": 90000.0})
    assert abs(sum(t[1] for t in table) - 1.0) < 1e-12
    assert [t[1] for t in table] == sorted((t[1] for t in table), reverse=True)
    eq, rl = opy.equal_length_test(rows)
    assert eq and rl == 150


def test_path_cov_ratio_is_the_sequential_f32_sum_not_the_rounded_exact_sum():
    """profile.rs:2714-2729: `RowDVector<f32> * DMatrix<f32>` accumulates in f32, node by node.  Once a running sum
    passes 2^24 the result is no longer (f32)sum_cov / (f32)sum_len formed from the exact integers - the oracle (and the
    host side of the product, pantax_b200.api.path_cov_ratio) must give the reference's value, not the better one."""
    rng = np.random.default_rng(5)
    n = 700_000
    lens = rng.integers(1, 64, n).astype(np.int64) | 1  # odd lengths: every partial sum needs its low bits
    cov = (lens - rng.integers(0, 2, n)).astype(np.int64)
    g = opy.Graph([int(x) for x in lens])
    g.paths = {"h1": list(range(n)), "h2": list(range(0, n, 3)) + [5, 5, 8]}
    got = opy.path_cov_ratio(g, cov.tolist())
    for h, p in enumerate([g.paths["h1"], g.paths["h2"]]):
        d = sorted(set(p))
        assert lens[d].sum() > 2 ** 24 or h == 1
        acc_c = acc_l = np.float32(0)
        for v in d:  # the naive loop
            acc_c = np.float32(acc_c + np.float32(cov[v]))
            acc_l = np.float32(acc_l + np.float32(lens[v]))
        assert got[h] == float(np.float32(acc_c / acc_l))
    exact = float(np.float32(cov.sum()) / np.float32(lens.sum()))
    assert got[0] != exact, "the test graph must be large enough for f32 accumulation to deviate from the exact sums"
    # f64 (the CBC variant, profile.rs:1952-1977) stays exact
    assert opy.path_cov_ratio(g, cov.tolist(), f32=False)[0] == cov.sum() / lens.sum()
