"""Host tail of strain profiling (pantax_b200/strain_tail.py: PAO model in CSR, the two filters, the abundance tables)
against the restatement of the reference in oracle/pantax_oracle.py.  CPU part: the model rows, the filters, the float
format.  GPU part: configs[1]-shaped input from the GAF all the way to strain_abundance.txt."""
import dataclasses
import os

import numpy as np
import pytest

from common import dataset_graphs, opy, py_graph, run_cpu_oracle, synth
from pantax_b200 import strain_tail as st


def _graph(rng, n=400, haps=5):
    lens = rng.integers(1, 60, n).astype(np.int64)
    paths = [np.array([v for v in range(n) if rng.random() < 0.7] + [3, 3], dtype=np.uint64) for _ in range(haps)]
    names = [f"GCF_{i:09d}.1" for i in range(haps)]
    return lens, paths, names


def test_pao_model_in_csr_is_the_reference_row_problem():
    """profile.rs:2699-2813: same variables, bounds, objective and rows (in the order of the add_row calls) as the
    dense construction - the solver inputs are identical, which is what makes the abundances identical."""
    rng = np.random.default_rng(3)
    lens, paths, names = _graph(rng)
    depth = np.where(rng.random(len(lens)) < 0.6, rng.random(len(lens)) * 30.0, 0.0)
    g = py_graph(lens, paths, names)
    for possible, pinned in (([0, 2, 3], []), ([1, 4], [1]), ([0, 1, 2, 3, 4], [0, 3])):
        m = st.build_pao_model(paths, possible, depth, st.ProfilingArgs(minimization_min_cov=0.5), fixed_zero=pinned)
        c, A, lo, hi, lb, ub, integer = opy.highs_rows_dense(g, possible, depth.tolist(), 0.5, fixed_zero=pinned)
        np.testing.assert_array_equal(m.c, c)
        np.testing.assert_array_equal(m.dense(), A)
        np.testing.assert_array_equal(m.lo, lo)
        np.testing.assert_array_equal(m.hi, hi)
        np.testing.assert_array_equal(m.lb, lb)
        np.testing.assert_array_equal(m.ub, ub)
        np.testing.assert_array_equal(m.integrality, integer)
        x = st.solve_pao(m)
        assert len(x) == len(c) and all(abs(x[i]) < 1e-9 for i in pinned)


def test_second_filter_and_abundance_constraint_match_the_restatement():
    rng = np.random.default_rng(9)
    for trial in range(300):
        H = int(rng.integers(1, 7))
        T = int(rng.integers(0, 3)) * H
        same = bool(rng.integers(0, 2))
        possible = sorted(rng.choice(H, size=int(rng.integers(1, H + 1)), replace=False).tolist())
        ms = []
        for h in range(H):
            ms.append(dict(otu="s", hap_id=f"h{h}", unique_trio_nodes_fraction=float(rng.choice([0.1, 0.5, 0.9, 1.0])),
                           frequencies_mean=float(rng.choice([0.0, 0.5, 3.0, 12.5])), path_cov_ratio=float(rng.choice([0.2, 0.9, 1.0])),
                           first_sol=float(rng.choice([0.0, 0.4, 3.1, 11.0, 40.0])), divergence=None, second_sol=None, is_rescue=None, total_cov_diff=None))
        opt = st.OptVar(otu="s", hap_metrics=[st.HapMetrics(**m) for m in ms], possible_paths_idx=list(possible), orign_n_haps=H,
                        hap2trio_nodes_m_size=T, same_path_flag=same)
        if (H == 1 or (T == 0 and same)) and 0 not in possible:
            continue
        st.second_filter_paths(opt, st.ProfilingArgs())
        so, keep = opy.second_filter_paths(ms, possible, H, T, same)
        assert opt.second_opt == so and opt.second_possible_paths_idx == keep
        for h in keep:  # the second solve would fill these
            opt.hap_metrics[h].second_sol = ms[h]["second_sol"] = float(rng.choice([0.0, 0.7, 5.0, 60.0]))
        cov = float(rng.choice([0.5, 4.0, 30.0]))
        st.abundace_constraint(cov, opt.hap_metrics)
        opy.abundace_constraint(cov, ms)
        for a, b in zip(opt.hap_metrics, ms):
            assert dataclasses.asdict(a) == b


def test_float_format_is_ryu_style():
    for v, s in [(1.0, "1.0"), (0.1, "0.1"), (150.0, "150.0"), (1e-5, "0.00001"), (1.5e-7, "1.5e-7"), (1e16, "1e16"), (1e15, "1000000000000000.0"),
                 (123456.789, "123456.789"), (1e-6, "1e-6"), (3e20, "3e20"), (-2.5, "-2.5"), (0.0, "0.0"), (None, "")]:
        assert st.fmt_f64(v) == s
    for v in np.random.default_rng(1).random(2000) * 10.0 ** np.random.default_rng(2).integers(-8, 20, 2000):
        assert float(st.fmt_f64(float(v))) == float(v)  # always round-trips


@pytest.mark.gpu
def test_gaf_to_strain_abundance_table_end_to_end(tmp_path):
    """configs[1]-shaped (one species, strain paths, short reads) from GAF text to the 11-column strain_abundance.txt:
    GPU for everything read-dependent, HiGHS for the ILP, and every intermediate compared with the restatement."""
    from pantax_b200 import api

    ds = synth.Dataset(61, [6000], [6])
    graphs = dataset_graphs(ds)
    lens, paths, names = graphs[0]
    gaf = ds.gaf(12, 0, 40000)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    ctx.upload_graph(0, lens, paths)
    ctx.commit_graphs()
    ctx.ingest_gaf(gaf, is_last=True)
    ctx.finalize()
    args = st.ProfilingArgs()
    metrics = st.optimize_otu(ctx, 0, ds.ranges()[0][0], lens, paths, names, args)
    # --- restatement on the CPU oracle's numbers
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    bases = o.node_bases(0)
    depth = (bases / lens).tolist()
    _k, tlen, owner = o.trio_table(0)
    tdepth = (o.trio_bases(0) / np.maximum(tlen, 1)).tolist()
    g = py_graph(lens, paths, names)
    possible, om, _same = opy.first_filter_paths(g, owner.tolist(), tdepth, depth, fr=0.3)
    assert [i for i, m in enumerate(metrics) if m.frequencies_mean is not None and i in possible] == possible
    ratio = opy.path_cov_ratio(g, o.node_cov(0).tolist())
    for i in possible:
        assert metrics[i].path_cov_ratio == ratio[i]
        assert metrics[i].unique_trio_nodes_fraction == om[i]["unique_trio_nodes_fraction"]
        assert metrics[i].frequencies_mean == pytest.approx(om[i]["frequencies_mean"], rel=1e-12)
    c, A, lo, hi, lb, ub, integer = opy.highs_rows_dense(g, possible, depth)
    m = st.build_pao_model(paths, possible, np.array(depth), args)
    np.testing.assert_array_equal(m.dense(), A)
    np.testing.assert_array_equal(m.hi, hi)
    # --- species coverage (species_abundance.txt: bases / mean genome length) and the final table
    counts = ctx.species_counts()
    species_cov = float(counts[0][1]) / float(int(lens.sum()) // len(paths))
    st.abundace_constraint(species_cov, metrics)
    info = [(f"G{i}", f"T{i}", ds.ranges()[0][0], "org", f"/db/{n}_genomic.fna") for i, n in enumerate(names)]
    rows = st.abundance_est(args, metrics, info, str(tmp_path / "strain_abundance.txt"), str(tmp_path / "ori_strain_abundance.txt"))
    txt = open(tmp_path / "strain_abundance.txt").read().split("\n")
    assert txt[0].split("\t") == ["species_taxid", "strain_taxid", "genome_ID", "predicted_coverage", "predicted_abundance", "path_base_cov",
                                  "unique_trio_fraction", "uniq_trio_cov_mean", "first_sol", "strain_cov_diff", "total_cov_diff"]
    assert len(rows) >= 1 and len(txt) == len(rows) + 2
    ab = [float(r[4]) for r in rows]
    assert ab == sorted(ab, reverse=True) and abs(sum(ab) - 1.0) < 1e-9
    assert all(r[2].startswith("G") and r[1].startswith("T") for r in rows)


def test_rust_round_is_exact_at_the_half_boundaries():
    assert st.rust_round(0.49999999999999994) == 0.0  # floor(x + 0.5) would say 1.0
    assert st.rust_round(0.5) == 1.0 and st.rust_round(-0.5) == -1.0 and st.rust_round(2.5) == 3.0 and st.rust_round(-2.5) == -3.0
    assert st.rust_round(4503599627370497.0) == 4503599627370497.0  # 2^52 + 1: x + 0.5 is not representable
    assert opy.round_half_away(0.49999999999999994) == 0.0 and opy.round_half_away(-1.5) == -2.0


def test_abundance_est_left_join_repeats_a_hap_with_two_genome_rows(tmp_path):
    """profile.rs:3176-3183: the left join on hap_id yields one row per matching genomes_info line; the abundance sum and the
    group size are taken over the joined rows."""
    args = st.ProfilingArgs()
    ms = []
    for i, sol in enumerate((3.0, 1.0)):
        m = st.HapMetrics(otu="562", hap_id=f"GCF_{i}", second_sol=sol, first_sol=sol, total_cov_diff=0.1)
        ms.append(m)
    info = [("G0", "T0", "562", "o", "/x/GCF_0_genomic.fna"), ("G0b", "T0b", "562", "o", "/y/GCF_0_other.fna.gz"), ("G1", "T1", "562", "o", "/x/GCF_1.fa")]
    rows = st.abundance_est(args, ms, info, str(tmp_path / "s.txt"), str(tmp_path / "o.txt"))
    assert [r[2] for r in rows] == ["G0", "G0b", "G1"]
    assert [float(r[4]) for r in rows] == [3.0 / 7.0, 3.0 / 7.0, 1.0 / 7.0]
    ori = open(tmp_path / "o.txt").read().split("\n")
    assert len(ori) == 1 + 3 + 1


def test_float_columns_print_like_the_readme_examples():
    """The only output rows the reference publishes (README.md:341-354, written by polars' CSV writer): every float there must be
    reproduced digit for digit by fmt_f64 from the parsed value."""
    readme = ["0.5005489240249426", "6.723225501680235", "16.0", "0.39983790355261384", "0.9967217217217217", "1.0", "15.54", "0.01",
              "0.0010005002501250622"]
    for s in readme:
        assert st.fmt_f64(float(s)) == s
    header = "species_taxid\tstrain_taxid\tgenome_ID\tpredicted_coverage\tpredicted_abundance\tpath_base_cov\tunique_trio_fraction\tuniq_trio_cov_mean\tfirst_sol\tstrain_cov_diff\ttotal_cov_diff"
    import inspect
    assert all(c in inspect.getsource(st.abundance_est) for c in header.split("\t"))


def test_hap_id_of_the_reference_example_genomes():
    """profile.rs:3106-3145 on the `id` column of the reference's own example tables (example/example_genomes_info.txt,
    genomes_info.txt): Path::file_stem drops ONE extension, then the first two `_` pieces are kept."""
    cases = {"../genomes/GCF_002012065.1_ASM201206v1_genomic.fna": "GCF_002012065.1",
             "../genomes/GCF_006400955.1_ASM640095v1_genomic.fna.gz": "GCF_006400955.1",
             "../genomes/MGYG000002538_genomic.fna": "MGYG000002538_genomic",
             "/path/to/GCF_009730575.1_ASM973057v1_genomic.fna": "GCF_009730575.1",
             "plain.fa": "plain"}
    for path, hap in cases.items():
        assert st.hap_id_of_genome(path) == hap
