"""The C++ host driver (pantax_b200/pantax-gpu-profile) on a synthetic PanTax database directory: the reference's
file formats in (species_range.txt, species_genomes_stats.txt, bincode .bin graphs, a GFA fallback, a GAF), its
output tables out - compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from common import NASTY, dataset_graphs, opy, py_graph, run_cpu_oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pantax_b200", "pantax-gpu-profile")


def write_bin(path, nodes_len, paths, names):
    """bincode 1.3 default (little-endian, fixed ints, u64 lengths) of types.rs:51-55 Graph, keys in byte order."""
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(nodes_len)))
        f.write(np.asarray(nodes_len, dtype="<i8").tobytes())
        f.write(struct.pack("<Q", len(paths)))
        for n, p in sorted(zip(names, paths), key=lambda t: t[0].encode()):
            k = n.encode()
            f.write(struct.pack("<Q", len(k)) + k + struct.pack("<Q", len(p)))
            f.write(np.asarray(p, dtype="<u8").tobytes())


def write_gfa(path, nodes_len, paths, names):
    with open(path, "w") as f:
        f.write("H\tVN:Z:1.1\n")
        for i, l in enumerate(nodes_len):
            f.write(f"S\t{i + 1}\t{'A' * int(l)}\n")
        for n, p in zip(names, paths):
            f.write(f"W\t{n}\t0\tchr1\t0\t100\t" + "".join(f">{int(v) + 1}" for v in p) + "\n")


def make_db(tmp, ds, graphs, gaf):
    db = os.path.join(tmp, "db")
    os.makedirs(os.path.join(db, "species_graph_info"))
    os.makedirs(os.path.join(db, "species_gfa"))
    with open(os.path.join(db, "species_range.txt"), "w") as f:
        for s, (t, a, b) in enumerate(ds.ranges()):
            f.write(f"{t}\t{a}\t{b}\t{1 if len(graphs[s][1]) > 1 else 0}\n")
    lens = {}
    with open(os.path.join(db, "species_genomes_stats.txt"), "w") as f:
        for s, (t, _a, _b) in enumerate(ds.ranges()):
            lens[t] = float(int(graphs[s][0].sum()) // max(1, len(graphs[s][1]))) + 0.5
            f.write(f"{t}\t{lens[t]}\n")
    for s, (t, _a, _b) in enumerate(ds.ranges()):
        if s == 1:  # one species only as text GFA: exercises the fallback reader (profile.rs:2923-2927)
            write_gfa(os.path.join(db, "species_gfa", f"{t}.gfa"), *graphs[s])
        elif s == 2:  # compressed forms of the same bincode stream (zip.rs:248-262)
            from golden.make_graph_fixtures import lz4_frame
            write_bin(os.path.join(tmp, "g.bin"), *graphs[s])
            open(os.path.join(db, "species_graph_info", f"{t}.bin.lz4"), "wb").write(lz4_frame(open(os.path.join(tmp, "g.bin"), "rb").read()))
        elif s == 3:
            from golden.make_graph_fixtures import zstd_frame
            write_bin(os.path.join(tmp, "g.bin"), *graphs[s])
            open(os.path.join(db, "species_graph_info", f"{t}.bin.zst"), "wb").write(zstd_frame(open(os.path.join(tmp, "g.bin"), "rb").read()))
        else:
            write_bin(os.path.join(db, "species_graph_info", f"{t}.bin"), *graphs[s])
    gp = os.path.join(tmp, "gfa_mapped.gaf")
    with open(gp, "wb") as f:
        f.write(gaf)
    return db, gp, lens


def test_host_driver_help_runs_without_gpu():
    out = subprocess.run([BIN, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--strain" in out.stdout


@pytest.mark.gpu
def test_host_driver_end_to_end_against_oracle(tmp_path):
    ds = synth.Dataset(404, [30000, 9000, 4000, 2500], [6, 3, 1, 2])  # .bin, .gfa, .bin.lz4, .bin.zst
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(8, 0, 60000, NASTY)
    db, gp, lens = make_db(str(tmp_path), ds, graphs, gaf)
    wd = os.path.join(str(tmp_path), "wd")
    rep = os.path.join(wd, "reads_classification.tsv")
    os.makedirs(wd)
    r = subprocess.run([BIN, "--db", db, "--gaf", gp, "--wd", wd, "--species", "--strain", "-R", rep, "-a", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ranges = ds.ranges()
    rows = opy.rcls_profile(gaf, ranges)
    # reads_classification.tsv (profile.rs:3337-3351): read_id, mapq, species, read_len
    got = [l.split("\t") for l in open(rep).read().split("\n")[:-1]]
    assert len(got) == len(rows)
    for g, row in zip(got, rows):
        assert g[0].encode() == row.read_id and g[2] == row.species
        assert g[1] == ("" if row.mapq is None else str(row.mapq)) and g[3] == ("" if row.read_len is None else str(row.read_len))
    # species_abundance.txt (profile.rs:299-349)
    exp = opy.species_profiling(rows, lens, filtered=True)
    lines = open(os.path.join(wd, "species_abundance.txt")).read().split("\n")
    assert lines[0] == "species_taxid\tpredicted_abundance\tpredicted_coverage"
    tab = [l.split("\t") for l in lines[1:] if l]
    assert [t[0] for t in tab] == [e[0] for e in exp]
    for t, e in zip(tab, exp):
        assert float(t[1]) == pytest.approx(e[1], rel=1e-13) and float(t[2]) == e[2]
    # strain inputs: integers bit-exact, single divisions identical, first_filter_paths decisions identical
    o = run_cpu_oracle(ranges, graphs, gaf)
    for s, (t, _a, _b) in enumerate(ranges):
        nodes = [l.split("\t") for l in open(os.path.join(wd, "strain_inputs", f"{t}.nodes.tsv")).read().split("\n")[1:] if l]
        bases, cov = o.node_bases(s), o.node_cov(s)
        nz = np.nonzero(bases)[0]
        assert [int(n[0]) for n in nodes] == nz.tolist()
        for n in nodes:
            i = int(n[0])
            assert float(n[2]) == bases[i] / graphs[s][0][i] and int(n[3]) == cov[i]
        paths = [l.split("\t") for l in open(os.path.join(wd, "strain_inputs", f"{t}.paths.tsv")).read().split("\n")[1:] if l]
        U, nzc = o.hap_trio_counts(s)
        sc, sl = o.path_sums(s)
        assert [int(p[1]) for p in paths] == U.tolist() and [int(p[2]) for p in paths] == nzc.tolist()
        assert [int(p[6]) for p in paths] == sc.tolist() and [int(p[7]) for p in paths] == sl.tolist()
        _k, tlen, owner = o.trio_table(s)
        trio_ab = (o.trio_bases(s) / np.maximum(tlen, 1)).tolist()
        node_ab = (bases / graphs[s][0]).tolist()
        possible, metrics, _same = opy.first_filter_paths(py_graph(*graphs[s]), owner.tolist(), trio_ab, node_ab, fr=0.3)
        assert [h for h, p in enumerate(paths) if p[8] == "1"] == possible
        for h, p in enumerate(paths):
            m = metrics[h]
            if m["unique_trio_nodes_fraction"] is not None:
                assert float(p[3]) == m["unique_trio_nodes_fraction"]
            if m["frequencies_mean"] is not None and p[4]:
                assert float(p[4]) == pytest.approx(m["frequencies_mean"], rel=1e-12)
        # profile.rs:2714-2728: sequential f32 accumulation in node-index order (nalgebra gemv), not (f32)sum / (f32)sum
        ratios = opy.path_cov_ratio(py_graph(*graphs[s]), o.node_cov(s).tolist())
        for h, p in enumerate(paths):
            assert float(p[5]) == ratios[h]


@pytest.mark.gpu
def test_host_driver_strain_only_resume_and_stdin(tmp_path):
    """`--strain` alone (profile.rs:3365-3419): the species column comes from reads_classification.tsv and the species
    table from the existing species_abundance.txt; the GAF arrives on stdin (the aligner's stdout, alignment.rs:18-26).
    The strain inputs must be identical to those of the full run."""
    ds = synth.Dataset(405, [20000, 7000], [5, 2])
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(9, 0, 40000, NASTY)
    db, gp, _lens = make_db(str(tmp_path), ds, graphs, gaf)
    wd = os.path.join(str(tmp_path), "wd")
    os.makedirs(wd)
    rep = os.path.join(wd, "reads_classification.tsv")
    r = subprocess.run([BIN, "--db", db, "--gaf", gp, "--wd", wd, "--species", "--strain", "-R", rep, "-a", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    full = {f: open(os.path.join(wd, "strain_inputs", f)).read() for f in sorted(os.listdir(os.path.join(wd, "strain_inputs")))}
    sa = open(os.path.join(wd, "species_abundance.txt")).read()
    # the same run fed through a pipe in 1 MB pinned chunks (reader thread + 3 buffers; lines straddle the chunk ends): identical
    # reads_classification.tsv and species table, and the stream throughput is reported
    wd2 = os.path.join(str(tmp_path), "wd2")
    os.makedirs(wd2)
    rep2 = os.path.join(wd2, "reads_classification.tsv")
    r = subprocess.run([BIN, "--db", db, "--gaf", "-", "--wd", wd2, "--species", "--strain", "-R", rep2, "-a", "0", "--chunk-mb", "1"], input=gaf, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert b"GAF stream:" in r.stderr and b"stdin" in r.stderr
    assert open(rep2).read() == open(rep).read()
    assert open(os.path.join(wd2, "species_abundance.txt")).read() == sa
    assert {f: open(os.path.join(wd2, "strain_inputs", f)).read() for f in sorted(os.listdir(os.path.join(wd2, "strain_inputs")))} == full
    for f in full:
        os.remove(os.path.join(wd, "strain_inputs", f))
    r = subprocess.run([BIN, "--db", db, "--gaf", "-", "--wd", wd, "--strain", "-a", "0"], input=gaf, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert b"Species column of" in r.stderr
    again = {f: open(os.path.join(wd, "strain_inputs", f)).read() for f in sorted(os.listdir(os.path.join(wd, "strain_inputs")))}
    assert again == full
    assert open(os.path.join(wd, "species_abundance.txt")).read() == sa   # not rewritten
    # a binning file that sends every read of species 2 to "U": that species gets no coverage
    t2 = ds.ranges()[1][0]
    rows = [l.split("\t") for l in open(rep).read().split("\n")[:-1]]
    with open(rep, "w") as f:
        for x in rows:
            f.write("\t".join([x[0], x[1], "U" if x[2] == t2 else x[2], x[3]]) + "\n")
    r = subprocess.run([BIN, "--db", db, "--gaf", gp, "--wd", wd, "--strain", "-a", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nodes2 = open(os.path.join(wd, "strain_inputs", f"{t2}.nodes.tsv")).read().split("\n")[1:]
    assert [l for l in nodes2 if l] == []
    t1 = ds.ranges()[0][0]
    assert open(os.path.join(wd, "strain_inputs", f"{t1}.nodes.tsv")).read() == full[f"{t1}.nodes.tsv"]


@pytest.mark.parametrize("ext", ["bin", "bin.lz4", "bin.zst"])
def test_graph_readers_on_the_bincode_fixtures(ext):
    """zip.rs:236-262: `.bin` (bincode), `.bin.lz4` (bincode inside an LZ4 frame), `.bin.zst` (bincode inside a zstd frame).
    The fixtures were written by tests/golden/make_graph_fixtures.py - bincode laid out by hand with struct, independently of
    the C++ reader - and the reader must return exactly that graph (no GPU involved: --dump-graph)."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "graph_fixture.json")))
    out = subprocess.run([BIN, "--dump-graph", os.path.join(ROOT, "tests", "golden", "graph_fixture." + ext)], capture_output=True, text=True, check=True).stdout
    lines = out.split("\n")
    assert lines[0] == f"nodes {len(g['nodes_len'])}"
    assert [int(x) for x in lines[1].split()] == g["nodes_len"]
    got = {}
    i = 2
    while i < len(lines) and lines[i].startswith("path "):
        _p, name, n = lines[i].split(" ")
        got[name] = [int(x) for x in lines[i + 1].split()]
        assert len(got[name]) == int(n)
        i += 2
    assert list(got) == sorted(g["paths"], key=lambda s: s.encode())  # BTreeMap order
    assert got == g["paths"]


def test_host_gfa_reader_trims_lines_like_the_reference(tmp_path):
    """profile.rs:483, 497: `line.trim()` before the split - CRLF files, trailing blanks behind a sequence and a trailing tab behind
    a walk must not change node lengths or hide the walk (the fallback C++ reader, --dump-graph; the device parser has its own test)."""
    txt = ("H\tVN:Z:1.1\r\nS\t1\tACGT\r\nS\t2\tAC \r\nS\t3\tA\t*\r\nS\t4\tGGG\t\r\n"
           "W\thap#B\t0\tchr0\t0\t100\t>3>1<2\t\r\nP\tGCF_1.1#1#chr1\t4+,1-\t* \r\nW\thap#B\t0\tchr1\t0\t100\t<4")
    p = str(tmp_path / "g.gfa")
    open(p, "w", newline="").write(txt)
    g = opy.read_gfa(txt.replace("\r\n", "\n"))
    assert g.nodes_len == [4, 2, 1, 3] and dict(g.paths) == {"GCF_1.1": [3, 0], "hap#B": [2, 0, 1, 3]}
    out = subprocess.run([BIN, "--dump-graph", p], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "nodes 4" and out[1] == "4 2 1 3"
    assert out[2:6] == ["path GCF_1.1 2", "3 0", "path hap#B 4", "2 0 1 3"]


def test_zip_gfa_writes_the_reference_bin_and_range_row(tmp_path):
    """zip.rs:78-171, 316-327: species GFA -> <stem>.bin (bincode Graph; W walks starting with '<' and P paths whose first step
    ends with '-' stored reversed) + the range row `stem, min, max, is_pan`; no GPU involved."""
    import random

    from pantax_b200 import strain_tail as st

    rng = random.Random(5)
    for trial in range(30):
        n = rng.randrange(3, 15)
        lines = ["H\tVN:Z:1.1"]
        for i in range(n):
            lines.append(f"S\t{i + 1}\t{''.join(rng.choice('ACGT') for _ in range(rng.randrange(1, 9)))}")
        for h in range(rng.randrange(1, 5)):
            name = rng.choice(["hapA", "hapB", "GCF_1.1"])
            ids = [rng.randrange(1, n + 1) for _ in range(rng.randrange(1, 8))]
            if rng.random() < 0.5:
                lines.append(f"W\t{name}\t0\tchr{h}\t0\t100\t" + "".join(rng.choice("<>") + str(v) for v in ids))
            else:
                lines.append(f"P\t{name}#1#chr{h}\t" + ",".join(str(v) + rng.choice("+-") for v in ids) + "\t*")
        txt = "\n".join(lines) + "\n"
        d = tmp_path / f"t{trial}"
        d.mkdir()
        gfa = str(d / "562.gfa")
        open(gfa, "w").write(txt)
        rf = str(d / "species_range.txt")
        subprocess.run([BIN, "--zip-gfa", gfa, str(d), rf], check=True)
        g, mn, mx, is_pan = opy.read_and_zip_gfa(txt)
        assert open(rf).read() == f"562\t{mn}\t{mx}\t{is_pan}\n"
        lens, names, paths = st.read_bin_graph(str(d / "562.bin"))
        assert lens.tolist() == g.nodes_len and names == list(g.paths)
        assert [p.tolist() for p in paths] == [g.paths[k] for k in names]
        # the byte stream is the bincode layout the independent writer of this file produces
        write_bin(str(d / "want.bin"), g.nodes_len, [g.paths[k] for k in names], names)
        assert open(str(d / "want.bin"), "rb").read() == open(str(d / "562.bin"), "rb").read()
        # a second run keeps the .bin (zip.rs:182) and appends another row
        subprocess.run([BIN, "--zip-gfa", gfa, str(d), rf], check=True)
        assert open(rf).read().count("\n") == 2


def test_existing_tables_are_not_recomputed(tmp_path):
    """profile.rs:3333-3425: a stage whose output table exists in the working directory is skipped; with both tables present the run
    ends at once (no GPU is touched, so this runs anywhere)."""
    wd = str(tmp_path)
    open(os.path.join(wd, "species_abundance.txt"), "w").write("species_taxid\tpredicted_abundance\tpredicted_coverage\n")
    open(os.path.join(wd, "strain_abundance.txt"), "w").write("x\n")
    r = subprocess.run([BIN, "--db", wd, "--gaf", os.path.join(wd, "none.gaf"), "--wd", wd, "--species", "--strain"], capture_output=True, text=True)
    assert r.returncode == 0 and "both exist" in r.stderr
    r = subprocess.run([BIN, "--db", wd, "--gaf", os.path.join(wd, "none.gaf"), "--wd", wd, "--species"], capture_output=True, text=True)
    assert r.returncode == 0 and "Species profiling abundance file exists" in r.stderr
