"""the species' GFA parsed on the device (ptx_upload_graph_gfa) against the Python restatement of
profile.rs:466-545 read_gfa - same nodes_len, same paths (haplotypes in BTreeMap order, chromosomes of one haplotype
concatenated in file order), and the same results downstream (trio table, coverage) as uploading the parsed arrays."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import NASTY, dataset_graphs, run_cpu_oracle, synth  # noqa: E402
from oracle import pantax_oracle as opy  # noqa: E402


def gfa_text(nodes_len, paths, names, rng, style):
    """A GFA the way pggb/vg write it, with the things read_gfa has to cope with: header and L lines, tags behind the sequence,
    several chromosomes of one genome (same haplotype id), P lines (`id+,id-`, name up to '#') and W lines (`>id<id`), CRLF."""
    eol = "\r\n" if style == "crlf" else "\n"
    out = ["H\tVN:Z:1.1"]
    for i, l in enumerate(nodes_len):
        tag = "\tLN:i:%d" % l if i % 7 == 0 else ""
        out.append("S\t%d\t%s%s" % (i + 1, "ACGT"[i % 4] * int(l), tag))
    for i in range(0, len(nodes_len) - 1, 997):
        out.append("L\t%d\t+\t%d\t+\t0M" % (i + 1, i + 2))
    lines = []
    for n, p in zip(names, paths):
        p = [int(v) for v in p]
        cuts = sorted(set([0, len(p)] + [int(c) for c in rng.integers(1, max(len(p), 2), size=2)]))  # up to three "chromosomes"
        for k in range(len(cuts) - 1):
            seg = p[cuts[k]:cuts[k + 1]]
            if not seg:
                continue
            if style == "P":
                lines.append("P\t%s#1#chr%d\t%s\t*" % (n, k, ",".join("%d%s" % (v + 1, "+-"[(v + k) % 2]) for v in seg)))
            else:
                lines.append("W\t%s\t0\tchr%d\t0\t%d\t%s" % (n, k, len(seg), "".join("%s%d" % ("><"[(v + k) % 2], v + 1) for v in seg)))
    order = rng.permutation(len(lines)) if style == "P" else np.arange(len(lines))  # P: chromosomes of different genomes interleaved
    # keep the chromosomes of ONE haplotype in their order (concatenation order is file order)
    by_name = {}
    for ln in lines:
        by_name.setdefault(ln.split("\t")[1].split("#")[0], []).append(ln)
    picked = []
    taken = {k: 0 for k in by_name}
    for j in order:
        name = lines[int(j)].split("\t")[1].split("#")[0]
        picked.append(by_name[name][taken[name]])
        taken[name] += 1
    out += picked
    return (eol.join(out) + eol).encode()


@pytest.mark.gpu
@pytest.mark.parametrize("style", ["W", "P", "crlf"])
def test_gfa_parsed_on_the_device_equals_read_gfa(style):
    from gpu_common import assert_gpu_matches_oracle
    from pantax_b200 import api

    rng = np.random.default_rng(5)
    ds = synth.Dataset(97, [30000, 7000, 1200], [6, 3, 1])
    graphs = dataset_graphs(ds)
    gfas = [gfa_text(g[0], g[1], g[2], rng, style) for g in graphs]
    parsed = []
    for s, text in enumerate(gfas):
        g = opy.read_gfa(text.decode(), 0)
        assert g.nodes_len == [int(x) for x in graphs[s][0]]
        sp = g.sorted_paths()
        parsed.append((np.asarray(g.nodes_len, dtype=np.int64), [np.asarray(p, dtype=np.uint64) for _n, p in sp], [n for n, _p in sp]))
        # (the generator only re-cuts the synthetic paths into chromosomes: the concatenation gives them back)
        assert [list(map(int, p)) for p in parsed[-1][1]] == [list(map(int, p)) for _n, p in sorted(zip(graphs[s][2], graphs[s][1]), key=lambda kv: kv[0].encode())]
    gaf = ds.gaf(11, 0, 30000, NASTY)
    o = run_cpu_oracle(ds.ranges(), parsed, gaf)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    for s, text in enumerate(gfas):
        ctx.upload_graph_gfa(s, text)
    for s in range(len(gfas)):  # the Graph itself: nodes_len, paths, haplotype ids
        nl, paths, names = ctx.species_graph(s)
        np.testing.assert_array_equal(nl, parsed[s][0])
        assert names == parsed[s][2]
        assert len(paths) == len(parsed[s][1])
        for a, b in zip(paths, parsed[s][1]):
            np.testing.assert_array_equal(a, b)
    ctx.commit_graphs()
    for s in range(len(gfas)):
        assert ctx.n_nodes(s) == len(parsed[s][0]) and ctx.n_paths(s) == len(parsed[s][1])
    ctx.ingest_gaf(gaf, is_last=True)
    ctx.finalize()
    assert_gpu_matches_oracle(ctx, o, parsed)
    ctx.close()


@pytest.mark.gpu
def test_gfa_errors_of_the_reference_are_error_codes():
    from pantax_b200 import api
    from pantax_b200._lib import PantaxGpuError

    def run(text, n_nodes):
        ctx = api.PantaxGpu(0)
        ctx.set_ranges([("1", 1, n_nodes)])
        try:
            ctx.upload_graph_gfa(0, text)
        finally:
            ctx.close()

    ok = b"S\t1\tAC\nS\t2\tG\nS\t3\tTTT\nP\tg1#1\t1+,2+,3-\t*\n"
    run(ok, 3)
    with pytest.raises(PantaxGpuError) as e:  # profile.rs:489
        run(b"S\t1\tAC\nS\t3\tG\nS\t2\tTTT\n", 3)
    assert e.value.name == "PTX_E_NODE_ORDER"
    with pytest.raises(PantaxGpuError) as e:  # profile.rs:494
        run(b"S\t1\tAC\nS\t2\t\tLN:i:0\nS\t3\tT\n", 3)
    assert e.value.name == "PTX_E_ZERO_LEN"
    with pytest.raises(PantaxGpuError) as e:  # fewer S lines than the species range has ids
        run(b"S\t1\tAC\nS\t2\tG\n", 3)
    assert e.value.name == "PTX_E_NVERT_MISMATCH"
    with pytest.raises(PantaxGpuError) as e:  # a path step outside the graph
        run(b"S\t1\tAC\nS\t2\tG\nS\t3\tT\nW\th\t0\tc\t0\t3\t>1>2>9\n", 3)
    assert e.value.name == "PTX_E_INVALID"
