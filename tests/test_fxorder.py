"""The reference's trio numbering (FxHashSet iteration order, profile.rs:659-716): the product's host-side emulation
(ptx_trio_ref_order, pantax_b200/csrc/ptx_fxorder.h) against the structural restatement in oracle/fx_hashset.py.
Host-only code - no GPU needed."""
import random

import numpy as np
import pytest

from oracle import fx_hashset as fx
from pantax_b200 import api, strain_tail
from pantax_b200._lib import PantaxGpuError


def unique_table_in_library_order(paths):
    """What ptx_trio_table returns: unique trios ordered by (owner hap, window position)."""
    count = {}
    for p in paths:
        for t in fx.canonical_windows(list(p)):
            count[t] = count.get(t, 0) + 1
    keys = []
    for p in paths:
        for t in fx.canonical_windows(list(p)):
            if count[t] == 1:
                keys.append(t)
    return np.array(keys, dtype=np.uint64).reshape(-1, 3)


def random_paths(rng, n_nodes, n_paths, mean_len):
    """Strain-like paths: a shared backbone with per-path skips, repeats and a few inversions."""
    paths = []
    for _ in range(n_paths):
        ln = max(0, int(rng.gauss(mean_len, mean_len / 3)))
        start = rng.randrange(max(1, n_nodes - ln)) if n_nodes > ln else 0
        p, v = [], start
        while len(p) < ln and v < n_nodes:
            if rng.random() > 0.1:
                p.append(v)
            if rng.random() < 0.02 and p:
                p.append(p[rng.randrange(len(p))])
            v += 1 if rng.random() < 0.9 else 2
        if rng.random() < 0.3:
            p.reverse()
        paths.append(p)
    return paths


def test_fx_hash_is_the_documented_recurrence():
    K = 0x517CC1B727220A95
    assert fx.fx_hash_words([0]) == 0
    assert fx.fx_hash_words([1]) == K
    # two words by hand: rotl(K, 5) ^ 2, times K
    r = ((K << 5) | (K >> 59)) & fx.MASK64
    assert fx.fx_hash_words([1, 2]) == ((r ^ 2) * K) & fx.MASK64


def test_table_geometry():
    assert [fx.capacity_to_buckets(c) for c in (1, 3, 4, 7, 8, 14, 15, 28, 29, 56, 57)] == [4, 4, 8, 8, 16, 16, 32, 32, 64, 64, 128]
    assert [fx.bucket_mask_to_capacity(b - 1) for b in (4, 8, 16, 32, 1024)] == [3, 7, 14, 28, 896]


def test_growth_follows_reserve_one_before_lookup():
    """A table that is exactly full grows on the next insert even when the key is already present (hashbrown >= 0.14)."""
    s = fx.FxHashSet()
    for i in range(3):
        s.insert((i, i, i))
    assert s.t.buckets == 4 and s.t.growth_left == 0
    s.insert((0, 0, 0))  # present
    assert s.t.buckets == 8
    assert sorted(s.into_iter()) == [(0, 0, 0), (1, 1, 1), (2, 2, 2)]


def test_extend_reserves_full_hint_when_empty_and_half_after():
    s = fx.FxHashSet()
    s.extend([(i, 0, 0) for i in range(100)])
    assert s.t.buckets == fx.capacity_to_buckets(100) == 128
    s.extend([(i, 1, 0) for i in range(60)])  # growth_left 12 < 30 -> resize(max(130, 113)) = 256 buckets
    assert s.t.buckets == 256


@pytest.mark.parametrize("seed,n_nodes,n_paths,mean_len", [(1, 30, 3, 8), (2, 200, 6, 60), (3, 2000, 12, 700), (4, 50, 1, 40),
                                                            (5, 6, 4, 5), (6, 20000, 20, 6000), (7, 3, 2, 3)])
def test_product_order_equals_structural_restatement(seed, n_nodes, n_paths, mean_len):
    rng = random.Random(seed)
    paths = random_paths(rng, n_nodes, n_paths, mean_len)
    keys = unique_table_in_library_order(paths)
    _all, uniq = fx.reference_trio_numbering(paths)
    order = api.trio_ref_order([np.array(p, dtype=np.int64) for p in paths], keys)
    assert len(order) == len(uniq) == len(keys)
    assert sorted(order.tolist()) == list(range(len(keys)))
    got = [tuple(int(x) for x in keys[i]) for i in order]
    assert got == uniq


def test_small_tables_and_degenerate_paths():
    for paths in ([[0, 1, 2]], [[0, 1]], [[], [5, 4, 3, 2, 1, 0]], [[0, 1, 2, 3], [0, 1, 2, 3]], [[1, 1, 1, 1, 1]],
                  [[0, 1, 2], [2, 1, 0], [7, 8, 9, 10]]):
        keys = unique_table_in_library_order(paths)
        _all, uniq = fx.reference_trio_numbering(paths)
        order = api.trio_ref_order([np.array(p, dtype=np.int64) for p in paths], keys)
        assert [tuple(int(x) for x in keys[i]) for i in order] == uniq


def test_mismatching_table_is_an_error():
    paths = [np.array([0, 1, 2, 3, 4])]
    keys = unique_table_in_library_order(paths)
    bad = keys.copy()
    bad[0, 1] += 100
    with pytest.raises(PantaxGpuError):
        api.trio_ref_order(paths, bad)
    with pytest.raises(PantaxGpuError):
        api.trio_ref_order(paths, keys[:-1])


def test_first_filter_sums_in_reference_order():
    """frequencies_mean adds a hap's trio abundances in the reference's trio order (profile.rs:1123-1146): with `trio_order` the
    f64 result equals a sequential sum in that order, bit for bit."""
    rng = random.Random(11)
    paths = random_paths(rng, 3000, 5, 1500)
    keys = unique_table_in_library_order(paths)
    T = len(keys)
    assert T > 200
    count = {}
    owner = np.zeros(T, dtype=np.uint32)
    pos = {tuple(int(x) for x in k): i for i, k in enumerate(keys)}
    for h, p in enumerate(paths):
        for t in fx.canonical_windows(p):
            if t in pos:
                owner[pos[t]] = h
    nprng = np.random.default_rng(5)
    depth = nprng.lognormal(1.0, 1.0, T) * (nprng.random(T) > 0.2)
    order = api.trio_ref_order([np.array(p) for p in paths], keys)
    args = strain_tail.ProfilingArgs()
    names = [f"h{i}" for i in range(len(paths))]

    def run(trio_order):
        opt = strain_tail.OptVar(otu="x", hap_metrics=[strain_tail.HapMetrics() for _ in names])
        strain_tail.first_filter_paths(opt, names, paths, owner, depth, np.zeros(3000), args, trio_order=trio_order)
        return opt

    a, b = run(order), run(None)
    assert a.possible_paths_idx == b.possible_paths_idx
    differs = 0
    for h in range(len(paths)):
        ma, mb = a.hap_metrics[h], b.hap_metrics[h]
        assert ma.unique_trio_nodes_fraction == mb.unique_trio_nodes_fraction
        if ma.frequencies_mean is None:
            continue
        vals = [float(depth[i]) for i in order if owner[i] == h and depth[i] > 0.0]
        kept = strain_tail.zscore_filter(vals, 3.0)
        s = 0.0
        for v in kept:
            s += v
        assert ma.frequencies_mean == (s / len(kept) if kept else 0.0)
        assert abs(ma.frequencies_mean - mb.frequencies_mean) <= 1e-12 * max(1.0, abs(mb.frequencies_mean))
        differs += ma.frequencies_mean != mb.frequencies_mean
    assert differs > 0  # the order is visible in the last bits, which is why it is reproduced


def test_oracle_in_reference_order_equals_strain_tail_bit_for_bit():
    """The restatement numbered like the reference (trio_nodes_info_reference_order) and the product's tail fed with
    ptx_trio_ref_order give the same frequencies_mean to the last bit."""
    from oracle import pantax_oracle as opy

    rng = random.Random(23)
    paths = random_paths(rng, 4000, 6, 2500)
    names = [f"hap{i:02d}" for i in range(len(paths))]
    lens = [rng.randrange(1, 60) for _ in range(4000)]
    g = opy.Graph(nodes_len=lens, paths={n: list(p) for n, p in zip(names, paths)})
    trio_map, _tlen, owner_ref = opy.trio_nodes_info_reference_order(g)
    lib_map, _l2, owner_lib = opy.trio_nodes_info(g)
    assert set(trio_map) == set(lib_map)
    keys = np.array(sorted(lib_map, key=lib_map.get), dtype=np.uint64).reshape(-1, 3)
    np.testing.assert_array_equal(keys, unique_table_in_library_order(paths))
    nprng = np.random.default_rng(9)
    depth_by_key = {k: float(nprng.lognormal(0.5, 1.2)) * float(nprng.random() > 0.15) for k in lib_map}
    # restatement, reference numbering
    ref_keys = sorted(trio_map, key=trio_map.get)
    possible, om, _same = opy.first_filter_paths(g, owner_ref, [depth_by_key[k] for k in ref_keys], [0.0] * 4000, fr=0.3)
    # product tail: library numbering + permutation
    lib_keys = [tuple(int(x) for x in k) for k in keys]
    order = api.trio_ref_order([np.array(p) for p in paths], keys)
    assert [lib_keys[i] for i in order] == ref_keys
    opt = strain_tail.OptVar(otu="x", hap_metrics=[strain_tail.HapMetrics() for _ in names])
    strain_tail.first_filter_paths(opt, names, paths, np.array(owner_lib, dtype=np.uint32), np.array([depth_by_key[k] for k in lib_keys]),
                                   np.zeros(4000), strain_tail.ProfilingArgs(), trio_order=order)
    assert opt.possible_paths_idx == possible
    for h in possible:
        assert opt.hap_metrics[h].frequencies_mean == om[h]["frequencies_mean"]
        assert opt.hap_metrics[h].unique_trio_nodes_fraction == om[h]["unique_trio_nodes_fraction"]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_colliding_keys_walk_the_triangular_probe_sequence(seed):
    """Trios chosen so that a third of all keys start probing in the same 16-byte group of the final table: the chain runs over
    many groups (16, 32, 48, ... apart) and through every resize.  The simplified emulation and the structural restatement must
    still agree slot for slot."""
    rng = random.Random(seed)
    chosen = []
    while len(chosen) < 300:
        a, b, c = rng.randrange(1 << 20), rng.randrange(1 << 20), rng.randrange(1 << 20)
        if a > c:
            a, c = c, a
        if fx.fx_hash_words((a, b, c)) & 0x7FF < 16:
            chosen.append((a, b, c))
    paths = [[v for t in chosen[k::3] for v in t] for k in range(3)]  # the windows across two triples add ~600 ordinary keys
    keys = unique_table_in_library_order(paths)
    allt, uniq = fx.reference_trio_numbering(paths)
    s = fx.FxHashSet()
    for p in paths:
        s.extend(fx.canonical_windows(p))
    assert s.t.buckets in (1024, 2048)
    starts = [fx.fx_hash_words(k) & s.t.bucket_mask for k in allt]
    assert sum(1 for v in starts if v < 16) >= 300  # 300 keys for the 16 slots of one group
    order = api.trio_ref_order([np.array(p, dtype=np.int64) for p in paths], keys)
    assert [tuple(int(x) for x in keys[i]) for i in order] == uniq
