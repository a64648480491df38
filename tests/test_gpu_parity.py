"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle."""
import numpy as np
import pytest

from common import LABEL_U, NASTY, NASTY_DUP, dataset_graphs, opy, run_cpu_oracle, synth
from test_oracle import kat_inputs, load_kats

pytestmark = pytest.mark.gpu


def _api():
    from pantax_b200 import api
    return api


@pytest.mark.parametrize("case", load_kats(), ids=lambda c: c["name"])
@pytest.mark.parametrize("flow", ["fused", "late"])
def test_gpu_matches_hand_derived_kats(case, flow):
    from gpu_common import gpu_vs_oracle
    ranges, graphs, gaf = kat_inputs(case)
    ctx, o = gpu_vs_oracle(ranges, graphs, gaf, flow=flow)
    exp = case["expect"]
    idx = {r[0]: i for i, r in enumerate(ranges)}
    assert ctx.read_labels().tolist() == [idx.get(l, LABEL_U) for l in exp["labels"]]
    for name, _s, _e in ranges:
        if name in exp:
            s = idx[name]
            assert ctx.node_bases(s).tolist() == exp[name]["bases"]
            assert ctx.node_cov(s).tolist() == exp[name]["cov"]
            assert ctx.trio_bases(s).tolist() == exp[name]["trio_bases"]
            assert ctx.path_sums(s)[0].tolist() == exp[name]["path_sum_cov"]
            assert ctx.hap_trio_counts(s)[1].tolist() == exp[name]["hap_nz"]


@pytest.mark.parametrize("params,seed", [(synth.GafParams(), 1), (NASTY, 2), (NASTY_DUP, 3)])
@pytest.mark.parametrize("flow", ["fused", "late"])
def test_gpu_short_reads_multi_species(params, seed, flow):
    from gpu_common import gpu_vs_oracle
    ds = synth.Dataset(300 + seed, [30000, 8000, 12000, 400], [8, 1, 3, 2])
    gaf = ds.gaf(seed, 0, 60000, params)
    ctx, o = gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf, flow=flow)
    if params is NASTY_DUP:
        assert not ctx.ids_unique and o.mixed_dropped > 0


@pytest.mark.parametrize("flow", ["fused", "late"])
def test_gpu_many_chunks_both_flows_with_mixed_groups(flow):
    """Several ptx_ingest_gaf calls (chunks 2.. take the single-pass path) with duplicated ids whose groups span
    species: the keep-mask replay runs over every chunk's record table; in the "late" flow (the reference's own
    order: classify, then load graphs) the whole coverage is a replay."""
    from gpu_common import gpu_vs_oracle
    ds = synth.Dataset(333, [25000, 7000, 9000], [7, 2, 3])
    gaf = ds.gaf(11, 0, 50000, NASTY_DUP)
    cuts = [len(gaf) // 5, 2 * len(gaf) // 5, 3 * len(gaf) // 5, 4 * len(gaf) // 5]
    ctx, o = gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf, flow=flow, split=cuts)
    assert not ctx.ids_unique and o.mixed_dropped > 0


def test_gpu_long_reads_hifi_dialect():
    from gpu_common import gpu_vs_oracle
    ds = synth.Dataset(8, [40000, 25000, 9000], [4, 2, 3], backbone_mean=400)
    gaf = ds.gaf(6, 0, 3000, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.1))
    gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf)


def test_gpu_chunked_ingest_splits_lines_anywhere():
    from gpu_common import assert_gpu_matches_oracle, run_gpu
    ds = synth.Dataset(21, [20000, 5000], [6, 2])
    gaf = ds.gaf(2, 0, 20000, NASTY)
    graphs = dataset_graphs(ds)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    rng = np.random.default_rng(5)
    for n_cuts in (1, 7, 40):
        cuts = rng.integers(1, len(gaf) - 1, size=n_cuts).tolist()
        ctx = run_gpu(ds.ranges(), graphs, gaf, split=cuts)
        assert_gpu_matches_oracle(ctx, o, graphs)
    # a chunk with no newline at all, and a final line without '\n'
    g2 = gaf[:-1]
    o2 = run_cpu_oracle(ds.ranges(), graphs, g2)
    ctx = run_gpu(ds.ranges(), graphs, g2, split=[10, 20, 30, len(g2) - 5])
    assert_gpu_matches_oracle(ctx, o2, graphs)


@pytest.mark.parametrize("single_pass", [True, False])
def test_gpu_single_pass_chunks_estimates_and_redo(single_pass, monkeypatch):
    """After the first chunk of a ctx the library ingests in ONE pass over the text: tile size and record-table
    capacity come from the line statistics seen so far, rows are numbered per tile.  A chunk that breaks the
    estimate (short lines after long ones; tiles with more than 1024 lines) is abandoned on the device and redone
    with the count pass - results, row order of the labels and the equal-length rule must not change."""
    from gpu_common import assert_gpu_matches_oracle
    api = _api()
    if not single_pass:
        monkeypatch.setenv("PTX_NO_SINGLE_PASS", "1")
    ds = synth.Dataset(77, [30000, 9000], [6, 3], backbone_mean=200)
    graphs = dataset_graphs(ds)
    long_part = ds.gaf(5, 0, 600, synth.GafParams(long_reads=True, id_pair_suffix=False))
    short_part = ds.gaf(6, 0, 30000, NASTY)
    # tiny lines: every 4 KB holds hundreds of (mostly invalid-for-coverage) rows -> more than 1024 lines per tile
    tiny = b"".join(b"t%d\t1\n" % i for i in range(40000))
    comment = b"@CO\tcomment line\n" * 50
    parts = [long_part, short_part, comment + tiny, short_part[: len(short_part) // 3], long_part[: len(long_part) // 2]]
    gaf = b"".join(parts)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    for s, g in enumerate(graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    for i, part in enumerate(parts):   # one ptx_ingest_gaf call per part: parts 2.. are single-pass candidates
        ctx.ingest_gaf(part, is_last=(i == len(parts) - 1))
    assert ctx.num_records == o.n_records       # resolves the in-flight counts before finalize
    ctx.finalize()
    assert_gpu_matches_oracle(ctx, o, graphs)
    eq, rl = ctx.equal_length()
    rows = opy.rcls_profile(gaf, ds.ranges())
    assert (eq, rl if eq else None) == opy.equal_length_test(rows)
    # the same ctx again after reset: now even the first chunk is single-pass (statistics persist)
    ctx.reset()
    ctx.ingest_gaf(short_part, is_last=True)
    np.testing.assert_array_equal(ctx.read_labels(), run_cpu_oracle(ds.ranges(), graphs, short_part).labels())
    ctx.finalize()
    assert_gpu_matches_oracle(ctx, run_cpu_oracle(ds.ranges(), graphs, short_part), graphs)


def test_gpu_steady_stream_of_chunks_does_not_reallocate():
    """A cudaFree/cudaMalloc in the middle of a stream of chunks synchronises the device (it cost 5-50 ms per
    10 M-record step when the single-pass capacity estimate moved by a fraction of a percent): after two warm-up
    cycles, reset + ingest in pieces must not allocate record tables any more."""
    api = _api()
    ds = synth.Dataset(81, [20000], [6])
    gaf = ds.gaf(2, 0, 60000)
    graphs = dataset_graphs(ds)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    ctx.upload_graph(0, graphs[0][0], graphs[0][1])
    ctx.commit_graphs()
    cuts = np.linspace(0, len(gaf), 8).astype(int)
    allocs = []
    for cycle in range(5):
        ctx.reset()
        for i in range(7):
            ctx.ingest_gaf(gaf[cuts[i]:cuts[i + 1]], is_last=(i == 6))
        ctx.finalize()
        allocs.append(ctx.stats()["table_allocs"])
    assert allocs[2] == allocs[3] == allocs[4], allocs
    from gpu_common import assert_gpu_matches_oracle
    assert_gpu_matches_oracle(ctx, run_cpu_oracle(ds.ranges(), graphs, gaf), graphs)


def test_gpu_partial_graphs_and_species_only():
    from gpu_common import assert_gpu_matches_oracle, run_gpu
    api = _api()
    ds = synth.Dataset(31, [10000, 5000, 3000], [5, 2, 1])
    gaf = ds.gaf(3, 0, 20000, NASTY)
    graphs = dataset_graphs(ds)
    graphs[1] = None
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    ctx = run_gpu(ds.ranges(), graphs, gaf)
    assert_gpu_matches_oracle(ctx, o, graphs)
    with pytest.raises(api.PantaxGpuError) as e:
        ctx.node_bases(1)
    assert e.value.name == "PTX_E_NO_GRAPH"
    # species-level only: no graph at all
    ctx = run_gpu(ds.ranges(), [None, None, None], gaf)
    np.testing.assert_array_equal(ctx.species_counts(), o.species_counts())
    np.testing.assert_array_equal(ctx.read_labels(), o.labels())
    eq, rl = ctx.equal_length()
    rows = opy.rcls_profile(gaf, ds.ranges())
    assert (eq, rl if eq else None) == opy.equal_length_test(rows)


def test_gpu_equal_length_rule_long_reads():
    from gpu_common import run_gpu
    ds = synth.Dataset(8, [40000, 25000], [4, 2], backbone_mean=400)
    gaf = ds.gaf(6, 0, 1500, synth.GafParams(long_reads=True, id_pair_suffix=False))
    ctx = run_gpu(ds.ranges(), [None, None], gaf)
    eq, _ = ctx.equal_length()
    assert eq is False


@pytest.mark.parametrize("long_mode", [None, "1"])
def test_gpu_dialect_edges_and_start_beyond_node_error(long_mode, monkeypatch):
    from gpu_common import gpu_vs_oracle
    if long_mode:
        monkeypatch.setenv("PTX_LONG_MODE", long_mode)   # the same lines through the warp-cooperative long-line kernel
    from test_core_host import test_core_handcrafted_dialect_edges  # noqa: F401  (same inputs)
    ranges = [("a", 1, 50), ("b", 51, 80)]
    graphs = [(np.full(50, 7, dtype=np.int64), [np.arange(50, dtype=np.uint64), np.array([3, 2, 1, 2, 3, 9], dtype=np.uint64)], ["p", "q"]),
              (np.full(30, 5, dtype=np.int64), [np.arange(30, dtype=np.uint64)], ["r"])]
    lines = [
        b"r1\t50\t0\t50\t+\t>3>4>5\t21\t2\t18\t16\t16\t60\ttp:A:P",
        b"r2\t50\t0\t50\t+\t>4>3>2>3>4\t35\t1\t30\t29\t29\t60",
        b"r3\t50\t0\t50\t+\t<10<9\t14\t0\t14\t14\t14\t5\r",
        b"r4\t+50\t0\t50\t+\t>60>61\t10\t-1\t4\t4\t4\t60",
        b"r5\t5x\t0\t50\t+\t>7\t7\t1\t3\t2\t2\tabc",
        b"r6\t50\t0\t50\t+\t>8>9",
        b"", b"\r", b"@HD\tVN:1.0",
        b"r7\t50\t0\t50\t+\t>0000000000000000000012>13\t14\t0\t9\t9\t9\t60",
        b"r8\t50\t0\t50\t+\t>49>50>51\t19\t0\t19\t19\t19\t60",
        b"r9\t50\t0\t50\t+\tchr1_12\t7\t0\t5\t5\t5\t3",
        b"*\t50\t0\t50\t+\t>20>21>22\t21\t7\t20\t13\t13\t60",
    ]
    gaf_ok = b"\n".join(lines) + b"\n"
    gpu_vs_oracle(ranges, graphs, gaf_ok)
    gaf_bad = gaf_ok + b"r10\t50\t0\t50\t+\t>30>31\t14\t8\t20\t12\t12\t60"  # start 8 > len 7: profile.rs:854
    ctx, o = gpu_vs_oracle(ranges, graphs, gaf_bad)
    assert o.species_error(0) == 1


def test_gpu_supplied_species_column_strain_only_resume():
    """ptx_ingest_labels (profile.rs:3367-3385): labels come from reads_classification.tsv, not from the walk."""
    from gpu_common import assert_gpu_matches_oracle
    from test_oracle import _supplied_labels
    api = _api()
    ds = synth.Dataset(62, [20000, 6000, 9000], [6, 2, 3])
    gaf = ds.gaf(8, 0, 60000, NASTY_DUP)
    graphs = dataset_graphs(ds)
    lab = _supplied_labels(ds.ranges(), graphs, gaf, 3)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf, labels=lab)

    def run(labels, pieces):
        ctx = api.PantaxGpu(0)
        ctx.set_ranges(ds.ranges())
        for s, g in enumerate(graphs):
            ctx.upload_graph(s, g[0], g[1])
        ctx.commit_graphs()
        half = labels.size // 2
        ctx.ingest_labels(labels[:half])   # successive calls append
        ctx.ingest_labels(labels[half:])
        cuts = np.linspace(0, len(gaf), pieces + 1).astype(int)
        for i in range(pieces):
            ctx.ingest_gaf(gaf[cuts[i]:cuts[i + 1]], is_last=(i == pieces - 1))
        return ctx

    for pieces in (1, 3):
        ctx = run(lab, pieces)
        ctx.finalize()
        assert_gpu_matches_oracle(ctx, o, graphs)
    # a label whose species range does not contain the walk: row treated as "U", finalize reports it
    lab2 = _supplied_labels(ds.ranges(), graphs, gaf, 4, p_wrong=0.01)
    o2 = run_cpu_oracle(ds.ranges(), graphs, gaf, labels=lab2)
    assert o2.label_out_of_range
    ctx = run(lab2, 2)
    with pytest.raises(api.PantaxGpuError) as e:
        ctx.finalize()
    assert e.value.name == "PTX_E_RANGE"
    assert_gpu_matches_oracle(ctx, o2, graphs)
    # too few labels for the rows / labels after rows / bad species index
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    ctx.ingest_labels(lab[:10])
    with pytest.raises(api.PantaxGpuError) as e:
        ctx.ingest_gaf(gaf, is_last=True)
    assert e.value.name == "PTX_E_STATE"
    with pytest.raises(api.PantaxGpuError) as e:
        ctx.ingest_labels(np.array([7], dtype=np.uint32))
    assert e.value.name == "PTX_E_RANGE"
    ctx.reset()
    ctx.ingest_gaf(gaf, is_last=True)   # reset dropped the supplied labels: classifier again
    ctx.finalize()
    np.testing.assert_array_equal(ctx.read_labels(), run_cpu_oracle(ds.ranges(), graphs, gaf).labels())


def test_gpu_overlapping_ranges_use_first_match_in_file_order():
    from gpu_common import gpu_vs_oracle
    # rcls.rs:253-257 takes the FIRST matching row; with overlapping rows the binary search is not used
    ranges = [("wide", 1, 100), ("narrow", 10, 20), ("tail", 101, 140)]
    graphs = [(np.full(100, 9, dtype=np.int64), [np.arange(100, dtype=np.uint64)], ["p"]), None,
              (np.full(40, 3, dtype=np.int64), [np.arange(40, dtype=np.uint64)], ["q"])]
    lines = [b"x%d\t20\t0\t20\t+\t>%d>%d\t18\t1\t15\t14\t14\t60" % (i, a, a + 1) for i, a in enumerate([12, 15, 99, 100, 101, 120, 5])]
    ctx, o = gpu_vs_oracle(ranges, graphs, b"\n".join(lines) + b"\n")
    assert ctx.read_labels().tolist() == [0, 0, 0, LABEL_U, 2, 2, 0]


def test_gpu_device_resident_buffer_and_reset():
    import torch
    from gpu_common import assert_gpu_matches_oracle
    api = _api()
    ds = synth.Dataset(41, [50000], [10])
    gaf = ds.gaf(9, 0, 100000)
    graphs = dataset_graphs(ds)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    ctx.upload_graph(0, graphs[0][0], graphs[0][1])
    ctx.commit_graphs()
    for _rep in range(2):
        bid, dptr = ctx.gaf_buffer_alloc(len(gaf))
        src = torch.frombuffer(bytearray(gaf), dtype=torch.uint8).cuda()
        # device-to-device copy into the library's padded buffer
        import ctypes as C
        cudart = C.CDLL("libcudart.so")
        assert cudart.cudaMemcpy(C.c_void_p(dptr), C.c_void_p(src.data_ptr()), C.c_size_t(len(gaf)), 3) == 0
        ctx.ingest_gaf_device(bid, len(gaf))
        ctx.finalize()
        assert_gpu_matches_oracle(ctx, o, graphs)
        ctx.reset()


def test_gpu_medium_single_species_config2_shape():
    """BASELINE config 2 scaled 1/20: 50k nodes, 50 strain paths, 500k short-read records."""
    from gpu_common import gpu_vs_oracle
    ds = synth.Dataset(20261019, [50000], [50])
    gaf = ds.gaf(2, 0, 500000)
    ctx, o = gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf)
    assert ctx.n_trios(0) > 0 and ctx.trio_bases(0).sum() > 0


def test_gpu_result_is_invariant_under_record_permutation():
    from gpu_common import run_gpu
    ds = synth.Dataset(51, [20000, 6000], [6, 3])
    gaf = ds.gaf(4, 0, 50000, NASTY_DUP)
    lines = gaf.split(b"\n")[:-1]
    perm = np.random.default_rng(1).permutation(len(lines))
    gaf2 = b"\n".join(lines[i] for i in perm) + b"\n"
    graphs = dataset_graphs(ds)
    a = run_gpu(ds.ranges(), graphs, gaf)
    b = run_gpu(ds.ranges(), graphs, gaf2)
    np.testing.assert_array_equal(a.species_counts(), b.species_counts())
    for s in range(2):
        np.testing.assert_array_equal(a.node_bases(s), b.node_bases(s))
        np.testing.assert_array_equal(a.node_cov(s), b.node_cov(s))
        np.testing.assert_array_equal(a.trio_bases(s), b.trio_bases(s))


@pytest.mark.parametrize("rows,long_mode", [(1, "0"), (3, "0"), (8, "0"), (2, "1"), (8, "1")])
def test_gpu_every_tile_size_gives_the_same_answer(rows, long_mode, monkeypatch):
    """The ingest tile is rows*4 KB (chosen from the mean line length); force the extremes.  rows=1 makes
    tiles with ~37 records, rows=8 tiles with more records than threads (second pass).  long_mode forces the
    thread-per-record kernel ("0") or the warp-cooperative long-line kernel ("1") on both read types."""
    from gpu_common import gpu_vs_oracle
    monkeypatch.setenv("PTX_TILE_ROWS", str(rows))
    monkeypatch.setenv("PTX_LONG_MODE", long_mode)
    ds = synth.Dataset(61, [20000, 6000], [6, 3])
    gaf = ds.gaf(4, 0, 40000, NASTY_DUP)
    gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf)
    gl = ds.gaf(6, 0, 2000, synth.GafParams(long_reads=True, id_pair_suffix=False))
    gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gl)


@pytest.mark.parametrize("long_mode", ["0", "1", None])
def test_gpu_long_walks_window_edges(long_mode, monkeypatch):
    """Walk columns built to hit the edges of the cooperative decode: runs crossing the 128-byte windows, 1-9 digit
    ids, ids of 10-18 digits (64-bit scan), runs of more than 18 digits (dropped, rcls.rs:244), an empty column,
    "*", a line that ends after column 6 with CRLF, separators other than <>, and lines far longer than the staged
    window (parsed from global memory)."""
    from gpu_common import gpu_vs_oracle
    if long_mode is not None:
        monkeypatch.setenv("PTX_LONG_MODE", long_mode)
    n = 120000
    ranges = [("a", 1, n), ("b", n + 1, n + 5000)]
    rng = np.random.default_rng(5)
    paths = [np.arange(0, n, 7, dtype=np.uint64), np.arange(n - 1, 0, -11).astype(np.uint64), rng.integers(0, n, 5000).astype(np.uint64)]
    graphs = [(rng.integers(1, 30, n).astype(np.int64), paths, ["h1", "h2", "h3"]),
              (np.full(5000, 9, dtype=np.int64), [np.arange(5000, dtype=np.uint64)], ["k"])]

    def line(i, ids, c8=0, extra=b"\t60\t60\t60\ttp:A:P", seps=b"><"):
        walk = b"".join(bytes([seps[j % len(seps)]]) + str(int(v)).encode() for j, v in enumerate(ids))
        return b"lr%d\t15000\t0\t15000\t+\t%s\t%d\t%d\t%d%s" % (i, walk, 20 * len(ids), c8, 15 * len(ids) + 3, extra)

    lines = []
    for i in range(400):  # random walks: lengths 1..600 nodes, ids of 1-6 digits, with repeats now and then
        w = int(rng.integers(1, 600))
        mode = i % 4
        if mode == 0:
            ids = np.sort(rng.choice(n, w, replace=False)) + 1          # strictly increasing
        elif mode == 1:
            ids = np.sort(rng.choice(n, w, replace=False))[::-1] + 1    # strictly decreasing
        elif mode == 2:
            ids = rng.integers(1, 1000, w)                               # repeats: first-occurrence rule
        else:
            ids = rng.integers(1, 10 ** int(rng.integers(1, 6)), w)
        lines.append(line(i, ids, c8=int(rng.integers(0, 2))))
    lines.append(line(1000, [5, 123456789012, 7]))                       # 12-digit id: not in any range -> U
    lines.append(line(1001, [5, 6], extra=b""))                          # line ends after column 9
    lines.append(b"lr1002\t100\t0\t100\t+\t>5>6>7\r")                   # ends after column 6, CRLF
    lines.append(b"lr1003\t100\t0\t100\t+\t>5>6>7")                      # ends after column 6
    lines.append(b"lr1004\t100\t0\t100\t+\t*\t0\t0\t0\t0\t0\t60")
    lines.append(b"lr1005\t100\t0\t100\t+\t\t0\t0\t0\t0\t0\t60")           # empty column
    lines.append(b"lr1006\t100\t0\t100\t+\t>5>" + b"9" * 19 + b">6\t30\t0\t20\t0\t0\t60")   # 19-digit run dropped: walk 5,6
    lines.append(line(1007, [11, 12, 13, 14], seps=b"_x"))               # other separators
    lines.append(line(1008, rng.integers(1, n, 9000)))                   # ~60 KB line: beyond any staged window
    lines.append(line(1009, np.arange(n + 1, n + 4000)))                 # long monotone walk in species b
    lines.append(line(1010, [n, n + 1]))                                 # spans two species -> U
    for pad in range(0, 130, 9):                                         # shift the walk start over every byte phase
        lines.append(b"p%s\t100\t0\t100\t+\t%s\t900\t0\t800\t0\t0\t60" % (b"x" * pad, b"".join(b">%d" % v for v in range(100 + pad, 160 + pad))))
    order = rng.permutation(len(lines))
    gaf = b"\n".join(lines[i] for i in order) + b"\n"
    ctx, o = gpu_vs_oracle(ranges, graphs, gaf)
    assert ctx.species_counts()[0, 0] > 400 and o.species_error(0) == 0 and o.species_error(1) == 0


def test_gpu_long_read_filter_matches_oracle():
    """K10 / gaf_filter.rs:44-97: best (matches, identity) per read id, mapq > 20, span > 1000, one line per id."""
    from common import ocpu
    api = _api()
    ds = synth.Dataset(8, [40000, 25000], [4, 2], backbone_mean=400)
    gl = ds.gaf(6, 0, 4000, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.3))
    ctx = api.PantaxGpu(0)
    exp = opy.filter_max_alignment(gl)
    got = api.filter_max_alignment_mt(ctx, gl)
    assert got == exp and 0 < len(got) < gl.count(b"\n")
    assert ocpu.filter_gaf(gl) == exp
    # no trailing newline, and the filtered file is a fixed point
    assert ctx.filter_gaf(gl[:-1]) == exp
    again = b"\n".join(exp) + b"\n"
    assert ctx.filter_gaf(again) == exp


def test_gpu_long_read_filter_handcrafted_edges():
    api = _api()

    def line(rid, qs, qe, matches, mapq, ident, extra=b""):
        f = [rid, b"20000", qs, qe, b"+", b">1>2", b"30000", b"5", b"9000", matches, b"9000", mapq, b"NM:i:3", b"AS:f:1", b"dv:f:0.01",
             b"id:f:" + ident]
        return b"\t".join(f) + extra
    lines = [
        line(b"a", b"0", b"5000", b"4000", b"60", b"0.99"),          # a: best by matches is the next line
        line(b"a", b"0", b"5000", b"4500", b"60", b"0.90"),
        line(b"b", b"0", b"5000", b"4000", b"60", b"0.95"),          # b: tie on matches -> higher identity wins
        line(b"b", b"0", b"5000", b"4000", b"60", b"9.6e-1"),
        line(b"c", b"0", b"5000", b"4000", b"20", b"0.99"),          # c: best line fails mapq > 20 -> nothing for c
        line(b"c", b"0", b"5000", b"3000", b"60", b"0.99"),
        line(b"d", b"0", b"1000", b"900", b"60", b"0.99"),           # d: span 1000 is not > 1000
        line(b"e", b"0", b"5000", b"4000", b"60", b"0.99"),          # e: two identical best lines -> first in file order
        line(b"e", b"10", b"5010", b"4000", b"60", b"0.990"),
        line(b"f", b"0", b"5000", b"40x0", b"60", b"0.99"),          # f: unparsable matches -> line ignored entirely
        line(b"f", b"0", b"5000", b"10", b"60", b"0.5"),
        b"  " + line(b"g", b"0", b"5000", b"4000", b"60", b"1") + b"  ",   # trim(): leading/trailing blanks
        line(b"h", b"0", b"5000", b"4000", b"60", b"0.99", b"\textra:Z:tag:with:colons"),  # > 16 columns: still column 16
        b"short\tline",
        b"",
        line(b"i", b"-5", b"+5000", b"4000", b"+60", b"+.75"),       # signs; ".75"
        line(b"j", b"0", b"5000", b"4000", b"60", b"abc"),           # identity not a number -> ignored
    ]
    data = b"\n".join(lines) + b"\n"
    exp = opy.filter_max_alignment(data)
    ctx = api.PantaxGpu(0)
    got = ctx.filter_gaf(data)
    assert got == exp
    ids = [l.strip().split(b"\t")[0] for l in got]
    assert ids == [b"a", b"b", b"e", b"f", b"g", b"h", b"i"]
    assert got[1].endswith(b"9.6e-1") and got[2].split(b"\t")[2] == b"0"
    # an identity that cannot be represented exactly is refused, not guessed
    bad = line(b"k", b"0", b"5000", b"4000", b"60", b"0.12345678901234567890")
    with pytest.raises(api.PantaxGpuError) as e:
        ctx.filter_gaf(bad + b"\n")
    assert e.value.name == "PTX_E_UNSUPPORTED"


def test_gpu_full_size_config2_bit_exact_and_chunk_invariant():
    """BASELINE.json configs[1] at FULL size (1 M nodes, 50 strain paths, 10 M short-read records, 1.1 GB of GAF):
    every integer output equals the multithreaded oracle's bit for bit, and splitting the input into uneven
    host chunks (lines cut anywhere) does not change a single value."""
    import ctypes as C
    from gpu_common import assert_gpu_matches_oracle
    from common import ocpu
    api = _api()
    ds = synth.Dataset(20261017 + 2, [1_000_000], [50])
    graphs = dataset_graphs(ds)
    buf, nbytes = ds.gaf_raw(20261017 + 2, 0, 10_000_000)
    try:
        o = ocpu.CpuOracle(0)
        o.set_ranges(ds.ranges())
        o.set_graph(0, graphs[0][0], graphs[0][1])
        o.prepare_graphs()
        o.run(buf.value, nbytes)
        assert o.n_records == 10_000_000
        ctx = api.PantaxGpu(0)
        ctx.set_ranges(ds.ranges())
        ctx.upload_graph(0, graphs[0][0], graphs[0][1])
        ctx.commit_graphs()
        ctx.ingest_gaf(buf.value, nbytes, is_last=True)
        ctx.finalize()
        assert_gpu_matches_oracle(ctx, o, graphs)
        ref_bases, ref_cov, ref_trio = ctx.node_bases(0), ctx.node_cov(0), ctx.trio_bases(0)
        assert int(ref_bases.sum()) > 10_000_000 * 100
        # path_cov_ratio (profile.rs:2714-2729) with the reference's arithmetic: a sequential f32 accumulation in node-index
        # order.  These paths sum to > 2^24 bases, where that is NOT (f32)sum_cov / (f32)sum_len of the exact integers.
        ratio = api.path_cov_ratio(ctx, 0, graphs[0][1], graphs[0][0])
        sc, sl = ctx.path_sums(0)
        assert int(sl.min()) > 2 ** 24
        lens32, cov32 = graphs[0][0].astype(np.float32), ref_cov.astype(np.float32)
        for h in (0, 17, 49):
            acc_c = acc_l = np.float32(0)
            for v in np.unique(graphs[0][1][h]).tolist():  # the naive loop
                acc_c = np.float32(acc_c + cov32[v])
                acc_l = np.float32(acc_l + lens32[v])
            assert ratio[h] == float(np.float32(acc_c / acc_l))
        assert np.any(ratio != (sc.astype(np.float32) / sl.astype(np.float32)).astype(np.float64))
        np.testing.assert_array_equal(api.path_cov_ratio(ctx, 0, graphs[0][1], graphs[0][0], f32=False), sc / sl)
        # same bytes in three uneven chunks that split lines
        ctx.reset()
        cuts = [0, 333_333_337, 333_333_337 + 700_000_001, nbytes]
        for i in range(3):
            ctx.ingest_gaf(buf.value + cuts[i], cuts[i + 1] - cuts[i], is_last=(i == 2))
        ctx.finalize()
        np.testing.assert_array_equal(ctx.node_bases(0), ref_bases)
        np.testing.assert_array_equal(ctx.node_cov(0), ref_cov)
        np.testing.assert_array_equal(ctx.trio_bases(0), ref_trio)
        np.testing.assert_array_equal(ctx.species_counts(), o.species_counts())
    finally:
        synth.lib().synth_free(buf)


def test_gpu_config3_shape_hifi_reads_on_100_species():
    """BASELINE configs[2] shape, scaled: 100 species, 1 M nodes (mean node ~300 bp), HiFi reads (mean 15 kb,
    ~45-node walks, lines of ~400 B that often cross tile boundaries), secondary alignments (duplicate ids)."""
    from gpu_common import gpu_vs_oracle
    rng = np.random.default_rng(3)
    n_sp = 100
    nodes = rng.integers(5000, 15000, size=n_sp).tolist()
    haps = rng.integers(1, 8, size=n_sp).tolist()
    ds = synth.Dataset(20261017 + 3, nodes, haps, backbone_mean=400)
    gaf = ds.gaf(20261017 + 3, 0, 60000, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.05))
    ctx, o = gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf)
    assert not ctx.ids_unique


def test_gpu_config4_shape_short_reads_on_1000_species():
    """BASELINE configs[3] shape, scaled: 1,000 species (binary-search classification), 2 M nodes, 1 M short reads,
    graphs uploaded only for a third of the species (the rest is classified and counted but not covered)."""
    from gpu_common import gpu_vs_oracle
    rng = np.random.default_rng(4)
    n_sp = 1000
    nodes = rng.integers(500, 3500, size=n_sp).tolist()
    haps = rng.integers(1, 6, size=n_sp).tolist()
    ds = synth.Dataset(20261017 + 4, nodes, haps)
    gaf = ds.gaf(20261017 + 4, 0, 1_000_000, NASTY)
    graphs = dataset_graphs(ds)
    for s in range(n_sp):
        if s % 3:
            graphs[s] = None
    gpu_vs_oracle(ds.ranges(), graphs, gaf)


@pytest.mark.parametrize("variant", ["0", "1", "2", "3"])
@pytest.mark.parametrize("flow", ["fused", "late"])
def test_gpu_every_scatter_variant_gives_the_same_coverage(variant, flow, monkeypatch):
    """north_star stage 2: per-lane RED / warp-aggregated (match.any) / per-CTA shared-memory table / sort + segmented
    reduce must all produce the oracle's integers - on a small graph (reads of one warp and one CTA do share nodes), with
    duplicated ids across species (the keep-mask replay) and with the graphs committed before or after the ingest."""
    from gpu_common import gpu_vs_oracle
    monkeypatch.setenv("PTX_SCATTER", variant)
    ds = synth.Dataset(77, [900, 300, 5000], [5, 2, 3])
    gaf = ds.gaf(8, 0, 60000, NASTY_DUP)
    ctx, o = gpu_vs_oracle(ds.ranges(), dataset_graphs(ds), gaf, flow=flow, split=[len(gaf) // 3, len(gaf) // 2])
    assert not ctx.ids_unique


@pytest.mark.parametrize("long_mode", ["1", "0"])
def test_gpu_ten_digit_node_ids_take_the_exact_parser_without_extra_csr_slots(long_mode, monkeypatch):
    """Node ids >= 10^9 (10 digits; ptx_set_ranges allows them up to 2^32) are beyond the word-wide decoders of both ingest
    kernels: every record goes through the exact parser.  In the long-read kernel such a record must reuse the CSR slots it
    reserved from its line length instead of reserving again - a whole chunk of them used to outgrow the node buffer."""
    from gpu_common import gpu_vs_oracle
    monkeypatch.setenv("PTX_LONG_MODE", long_mode)
    rng = np.random.default_rng(11)
    base = 3_000_000_001
    n_nodes = 5000
    lens = rng.integers(1, 40, n_nodes).astype(np.int64)
    paths = [np.arange(n_nodes, dtype=np.uint64), np.arange(0, n_nodes, 2, dtype=np.uint64)]
    ranges = [("big", base, base + n_nodes - 1)]
    lines = []
    for i in range(30000):
        w = int(rng.integers(1, 60))
        s0 = int(rng.integers(0, n_nodes - w))
        ids = range(s0, s0 + w) if rng.random() < 0.5 else range(s0 + w - 1, s0 - 1, -1)
        walk = "".join((">" if rng.random() < 0.5 else "<") + str(base + v) for v in ids)
        plen = int(sum(lens[v] for v in ids))
        ps = int(rng.integers(0, lens[ids[0]] + 1))
        pe = int(rng.integers(ps, plen + 1))
        lines.append(f"r{i}\t{plen}\t0\t{plen}\t+\t{walk}\t{plen}\t{ps}\t{pe}\t{pe - ps}\t{pe - ps}\t60\ttp:A:P".encode())
    gaf = b"\n".join(lines) + b"\n"
    gpu_vs_oracle(ranges, [(lens, paths, ["h1", "h2"])], gaf, split=[len(gaf) // 2])


def test_gpu_text_of_resolved_chunks_is_released_and_reused():
    """Device memory must not grow with ~3 bytes per GAF byte: once a chunk's counts are in and the first 1000 non-U rows
    (ptx_equal_length, profile.rs:311-322) are known to lie in earlier chunks, its text buffer is handed back and reused by
    later pieces; the replay passes (mixed id groups, graphs committed after the ingest) run from the record tables."""
    api = _api()
    ds = synth.Dataset(83, [30000, 9000], [6, 3])
    gaf = ds.gaf(3, 0, 120000, NASTY_DUP)
    graphs = dataset_graphs(ds)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    cuts = np.linspace(0, len(gaf), 25).astype(int)
    for i in range(24):
        ctx.ingest_gaf(gaf[cuts[i]:cuts[i + 1]], is_last=(i == 23))
        if i % 6 == 5:
            assert ctx.num_records > 0  # waits for the chunks so far: their text can go back to the pool
    ctx.finalize()
    st = ctx.stats()
    assert st["text_buffers_released"] > 0 or st["text_bytes_resident"] < st["text_bytes"] // 2, st
    assert st["text_bytes_resident"] < st["text_bytes"], st
    eq, rl = ctx.equal_length()
    assert (eq, rl) == (True, 150)
    np.testing.assert_array_equal(ctx.read_labels(), o.labels())
    # graphs only now (the reference's order): the coverage pass replays the record tables - no text is needed
    for s, g in enumerate(graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    ctx.finalize()
    from gpu_common import assert_gpu_matches_oracle
    assert_gpu_matches_oracle(ctx, o, graphs)


def test_gpu_id_set_epochs_wrap_around():
    """The read-id set is not cleared between passes: its slots carry the epoch of the pass that wrote them (1..255) and a new pass
    only increments the epoch; the table is cleared for real when the epoch wraps.  270 passes over inputs whose duplicated ids form
    mixed-species groups (profile.rs:406-437) - alternating between two inputs, so that a stale slot of the other input mistaken for
    a live one would change the result - must each equal the oracle, before, at and after the wrap."""
    from gpu_common import assert_gpu_matches_oracle
    api = _api()
    ds = synth.Dataset(43, [6000, 2500, 900], [4, 2, 1])
    graphs = dataset_graphs(ds)
    gafs = [ds.gaf(21, 0, 6000, NASTY_DUP), ds.gaf(22, 3000, 9000, NASTY_DUP)]
    oracles = [run_cpu_oracle(ds.ranges(), graphs, g) for g in gafs]
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    for s, g in enumerate(graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    checked = 0
    for it in range(270):
        k = it & 1
        ctx.ingest_gaf(gafs[k], is_last=True)
        ctx.finalize()
        if it < 4 or 250 <= it <= 260 or it >= 266:
            assert_gpu_matches_oracle(ctx, oracles[k], graphs)
            checked += 1
        else:
            assert ctx.num_records == oracles[k].n_records
        ctx.reset()
    assert checked >= 18
    ctx.close()
