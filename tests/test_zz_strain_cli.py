"""File-level strain stage (`python -m pantax_b200.strain_tail --db DB --wd WD`): the tables pantax-gpu-profile leaves in
<wd>/strain_inputs -> strain_abundance.txt.  CPU part: the files are written here from the restatement's numbers in the
driver's formats (pantax_gpu_profile.cpp), the stage must give what the in-memory tail gives; the plain-.bin writer of the driver
is checked through --convert-graph.  GPU part (this file sorts last in the suite): the real driver, then the stage, against the in-process flow."""
import filecmp
import os
import subprocess
import sys

import numpy as np
import pytest

from common import dataset_graphs, opy, py_graph, run_cpu_oracle, synth
from pantax_b200 import strain_tail as st
from test_host_driver import BIN, make_db, write_bin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def f32_ratio(graph, cov):
    return opy.path_cov_ratio(graph, cov)


def write_strain_inputs(wd, taxid, lens, paths, names, o, s, fr=0.3):
    """What pantax_gpu_profile.cpp writes for one species (nodes.tsv, paths.tsv), from the restatement's numbers."""
    si = os.path.join(wd, "strain_inputs")
    os.makedirs(si, exist_ok=True)
    bases, cov = o.node_bases(s), o.node_cov(s)
    depth = bases / lens
    with open(os.path.join(si, f"{taxid}.nodes.tsv"), "w") as f:
        f.write("node\tlen\tdepth\tcovered_bases\n")
        for i in np.nonzero(depth > 0)[0]:
            f.write(f"{i}\t{int(lens[i])}\t{st.fmt_f64(float(depth[i]))}\t{int(cov[i])}\n")
    g = py_graph(lens, paths, names)
    _k, tlen, owner = o.trio_table(s)
    tdepth = (o.trio_bases(s) / np.maximum(tlen, 1)).tolist()
    possible, om, _same = opy.first_filter_paths(g, owner.tolist(), tdepth, depth.tolist(), fr=fr)
    U, nz = o.hap_trio_counts(s)
    sc, sl = o.path_sums(s)
    ratio = f32_ratio(g, cov.tolist())
    with open(os.path.join(si, f"{taxid}.paths.tsv"), "w") as f:
        f.write("hap_id\tunique_trios\tunique_trios_covered\tunique_trio_fraction\tuniq_trio_cov_mean\tpath_base_cov\tsum_cov\tsum_len\tpossible\n")
        for h, n in enumerate(names):
            frac = st.fmt_f64(om[h]["unique_trio_nodes_fraction"])
            mean = st.fmt_f64(om[h]["frequencies_mean"]) if h in possible else ""
            f.write(f"{n}\t{int(U[h])}\t{int(nz[h])}\t{frac}\t{mean}\t{st.fmt_f64(float(ratio[h]))}\t{int(sc[h])}\t{int(sl[h])}\t{1 if h in possible else 0}\n")
    return depth, possible, om, ratio


def in_memory_rows(args, species, cov_of, info, tmp):
    """The same tail without files: OptVar from the restatement's first filter, then the product's highs_opt / constraint / table."""
    metrics = []
    for taxid, lens, paths, names, depth, possible, om, ratio in species:
        if not np.any(depth > 0):
            continue
        H = len(names)
        opt = st.OptVar(otu=taxid, hap_metrics=[st.HapMetrics(otu=taxid, hap_id=n) for n in names])
        g = py_graph(lens, paths, names)
        trio_map, _tl, _ow = opy.trio_nodes_info(g)
        T = len(trio_map)
        opt.orign_n_haps, opt.hap2trio_nodes_m_size = H, H * T
        for h in range(H):
            opt.hap_metrics[h].unique_trio_nodes_fraction = om[h]["unique_trio_nodes_fraction"]
            if h in possible:
                opt.hap_metrics[h].frequencies_mean = om[h]["frequencies_mean"]
        opt.possible_paths_idx = list(possible)
        if H > 1 and T == 0:
            opt.same_path_flag = all(list(p) == list(paths[0]) for p in paths[1:])
        if opt.possible_paths_idx:
            st.highs_opt(opt, paths, depth, np.array(ratio), args)
        st.abundace_constraint(cov_of[taxid], opt.hap_metrics)
        metrics.extend(opt.hap_metrics)
    return st.abundance_est(args, metrics, info, os.path.join(tmp, "mem_strain_abundance.txt"), os.path.join(tmp, "mem_ori.txt"))


def make_genomes_info(db, ranges, graphs):
    info = []
    with open(os.path.join(db, "genomes_info.txt"), "w") as f:
        f.write("genome_ID\tstrain_taxid\tspecies_taxid\torganism_name\tid\n")
        for s, (t, _a, _b) in enumerate(ranges):
            for i, n in enumerate(graphs[s][2]):
                row = (f"G{s}_{i}", f"T{s}_{i}", t, f"org {s}", f"/genomes/{n}_genomic.fna")
                info.append(row)
                f.write("\t".join(row) + "\n")
    return info


def test_read_bin_graph_round_trip(tmp_path):
    lens = np.array([5, 1, 9, 300], dtype=np.int64)
    paths = [np.array([0, 1, 3], dtype=np.uint64), np.array([], dtype=np.uint64), np.array([3, 2, 1, 0, 0], dtype=np.uint64)]
    names = ["hapB", "hapA", "GCF_000001.1"]
    p = str(tmp_path / "g.bin")
    write_bin(p, lens, paths, names)
    gl, gn, gp = st.read_bin_graph(p)
    order = sorted(range(3), key=lambda i: names[i].encode())
    assert gn == [names[i] for i in order]
    np.testing.assert_array_equal(gl, lens)
    for a, i in zip(gp, order):
        np.testing.assert_array_equal(a, paths[i].astype(np.int64))
    open(p, "ab").write(b"x")
    with pytest.raises(ValueError):
        st.read_bin_graph(p)
    open(p, "wb").write(open(p, "rb").read()[:20])
    with pytest.raises(ValueError):
        st.read_bin_graph(p)


@pytest.mark.parametrize("ext", ["bin", "bin.lz4", "bin.zst"])
def test_driver_writes_the_plain_bincode_graph(tmp_path, ext):
    """--convert-graph: what the driver leaves as <wd>/strain_graphs/<taxid>.bin for a compressed / GFA species is the plain
    bincode stream of the same Graph - byte-identical to the independent fixture, and readable by the strain stage."""
    src = os.path.join(ROOT, "tests", "golden", "graph_fixture." + ext)
    out = str(tmp_path / "o.bin")
    subprocess.run([BIN, "--convert-graph", src, out], check=True)
    assert filecmp.cmp(out, os.path.join(ROOT, "tests", "golden", "graph_fixture.bin"), shallow=False)
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "graph_fixture.json")))
    lens, names, paths = st.read_bin_graph(out)
    assert lens.tolist() == g["nodes_len"] and names == sorted(g["paths"], key=lambda s: s.encode())
    assert [p.tolist() for p in paths] == [g["paths"][n] for n in names]


def test_strain_stage_from_files_equals_the_in_memory_tail(tmp_path):
    tmp = str(tmp_path)
    ds = synth.Dataset(77, [3000, 2500, 1500], [5, 1, 3])
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(5, 0, 30000)
    ranges = ds.ranges()
    db, _gp, lens_of = make_db(tmp, ds, graphs, gaf)
    info = make_genomes_info(db, ranges, graphs)
    wd = os.path.join(tmp, "wd")
    os.makedirs(wd)
    o = run_cpu_oracle(ranges, graphs, gaf)
    counts = o.species_counts()
    cov_of = {}
    with open(os.path.join(wd, "species_abundance.txt"), "w") as f:
        f.write("species_taxid\tpredicted_abundance\tpredicted_coverage\n")
        for s, (t, _a, _b) in enumerate(ranges):
            cov_of[t] = float(counts[s][1]) / lens_of[t]
            f.write(f"{t}\t{st.fmt_f64(1.0 / len(ranges))}\t{st.fmt_f64(cov_of[t])}\n")
    species = []
    for s, (t, _a, _b) in enumerate(ranges):
        lens, paths, names = graphs[s]
        byname = sorted(range(len(names)), key=lambda i: names[i].encode())
        assert byname == list(range(len(names)))  # the synthetic names are already in BTreeMap order
        depth, possible, om, ratio = write_strain_inputs(wd, t, lens, paths, names, o, s)
        species.append((t, lens, paths, names, depth, possible, om, ratio))
        if s != 0:  # species 1.. are GFA / lz4 / zst in make_db: the driver would leave their graph as a plain .bin in strain_graphs/
            os.makedirs(os.path.join(wd, "strain_graphs"), exist_ok=True)
            write_bin(os.path.join(wd, "strain_graphs", f"{t}.bin"), lens, paths, names)
    args = st.ProfilingArgs()
    ori = os.path.join(tmp, "ori.txt")
    rc = subprocess.run([sys.executable, "-m", "pantax_b200.strain_tail", "--db", db, "--wd", wd, "--ori", ori], cwd=ROOT, capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    want = in_memory_rows(args, species, cov_of, info, tmp)
    got = [l.split("\t") for l in open(os.path.join(wd, "strain_abundance.txt")).read().split("\n")[1:] if l]
    assert len(got) >= 2
    assert got == want
    assert open(ori).read() == open(os.path.join(tmp, "mem_ori.txt")).read()
    ab = [float(r[4]) for r in got]
    assert ab == sorted(ab, reverse=True) and abs(sum(ab) - 1.0) < 1e-9
    assert {r[0] for r in got} <= {t for t, _a, _b in ranges}


@pytest.mark.gpu
def test_zz_driver_then_strain_stage_on_the_gpu(tmp_path):
    """pantax-gpu-profile --species --strain, then the strain stage on its files == the in-process flow (optimize_otu over the
    C ABI) for every species the driver chose."""
    from pantax_b200 import api

    tmp = str(tmp_path)
    ds = synth.Dataset(78, [3000, 2500, 1500, 1200], [5, 2, 3, 4])
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(6, 0, 40000)
    ranges = ds.ranges()
    db, gp, lens_of = make_db(tmp, ds, graphs, gaf)
    info = make_genomes_info(db, ranges, graphs)
    wd = os.path.join(tmp, "wd")
    os.makedirs(wd)
    r = subprocess.run([BIN, "--db", db, "--gaf", gp, "--wd", wd, "--species", "--strain", "-a", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # species 1 (GFA), 2 (.bin.lz4), 3 (.bin.zst) get their graph as a plain .bin; species 0 is read from the db
    chosen = [t for t, _a, _b in ranges if os.path.exists(os.path.join(wd, "strain_inputs", f"{t}.paths.tsv"))]
    assert chosen
    for s, (t, _a, _b) in enumerate(ranges):
        if t in chosen:
            assert os.path.exists(os.path.join(wd, "strain_graphs", f"{t}.bin")) == (s != 0)
    ori = os.path.join(tmp, "ori.txt")
    args = st.ProfilingArgs()
    rows = st.run_strain_stage(db, wd, args, ori)
    # in-process flow over the C ABI
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ranges)
    for s, (lens, paths, _n) in enumerate(graphs):
        ctx.upload_graph(s, lens, paths)
    ctx.commit_graphs()
    ctx.ingest_gaf(gaf, is_last=True)
    ctx.finalize()
    cov_of = {r[0]: float(r[2]) for r in st._tsv(os.path.join(wd, "species_abundance.txt"))}
    metrics = []
    for s, (t, _a, _b) in enumerate(ranges):
        if t not in chosen:
            continue
        lens, paths, names = graphs[s]
        if not np.any(api.get_node_abundances(ctx, s)[0] > 0):
            continue
        m = st.optimize_otu(ctx, s, t, lens, paths, names, args)
        st.abundace_constraint(cov_of[t], m)
        metrics.extend(m)
    want = st.abundance_est(args, metrics, info, os.path.join(tmp, "mem.txt"), os.path.join(tmp, "mem_ori.txt"))
    assert [r[:3] for r in rows] == [r[:3] for r in want]
    for a, b in zip(rows, want):
        for x, y in zip(a[3:], b[3:]):
            assert (x == "" and y == "") or float(x) == pytest.approx(float(y), rel=1e-9, abs=1e-12)


@pytest.mark.gpu
def test_zz_filter_gaf_file_mode(tmp_path):
    """alignment.rs:171 at file level: `pantax-gpu-profile --filter-gaf gfa_mapped.gaf` writes gfa_mapped_filtered.gaf next to it with
    the lines gaf_filter::filter_max_alignment_mt keeps (CRLF input included)."""
    ds = synth.Dataset(8, [40000, 25000], [4, 2], backbone_mean=400)
    gl = ds.gaf(6, 0, 3000, synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.3))
    exp = opy.filter_max_alignment(gl)
    assert 0 < len(exp) < gl.count(b"\n")
    for name, data in (("gfa_mapped.gaf", gl), ("crlf.gaf", gl.replace(b"\n", b"\r\n"))):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        r = subprocess.run([BIN, "--filter-gaf", p], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = str(tmp_path / (name[:-4] + "_filtered.gaf"))
        assert out in r.stdout
        assert open(out, "rb").read() == b"".join(l + b"\n" for l in exp)


def test_profile_orchestrator_plan():
    """`python -m pantax_b200 profile`: filter (long reads, on request) -> pantax-gpu-profile -> strain stage, options routed to the
    stage that reads them, unknown ones handed to the driver."""
    from pantax_b200.__main__ import BIN as B, plan

    c = plan(["--db", "DB", "--gaf", "x/gfa_mapped.gaf", "--wd", "OUT", "--long-read", "--filter", "--fr", "0.5", "--fc", "0.4", "-a", "0", "--ds", "562,34"])
    assert c[0] == [B, "--filter-gaf", "x/gfa_mapped.gaf"]
    assert c[1][:8] == [B, "--db", "DB", "--gaf", "x/gfa_mapped_filtered.gaf", "--wd", "OUT", "--species"]
    assert "--strain" in c[1] and "--long-read" in c[1] and c[1][-4:] == ["-a", "0", "--ds", "562,34"] and "--fc" not in c[1]
    assert c[2][1:3] == ["-m", "pantax_b200.strain_tail"] and c[2][c[2].index("--fc") + 1] == "0.4" and c[2][c[2].index("--fr") + 1] == "0.5"
    c = plan(["--db", "DB", "--gaf", "-", "--species-only"])
    assert len(c) == 1 and "--strain" not in c[0] and c[0][c[0].index("--gaf") + 1] == "-"
    r = subprocess.run([sys.executable, "-m", "pantax_b200"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 2 and "profile --db" in r.stdout
