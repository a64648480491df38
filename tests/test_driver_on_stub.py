"""The real `pantax-gpu-profile` binary on a machine without a GPU: LD_LIBRARY_PATH points it at tests/stub_gpu.cpp, a stand-in
libpantax_gpu.so that answers the C ABI from the C++ restatement.  What is under test here is the DRIVER - the reference's file
formats in and out, streaming through pinned chunks and stdin, the strain-only resume, reads_classification.tsv, the species
table, strain_inputs / strain_graphs, the reference trio order in frequencies_mean, the long-read filter at file level, the
hand-off to the strain stage - by running the bodies of the `-m gpu` driver tests unchanged.  It says nothing about the kernels."""
import os
import subprocess

import numpy as np
import pytest

import test_host_driver as thd
import test_zz_strain_cli as tsc
from common import dataset_graphs, opy, run_cpu_oracle, synth
from pantax_b200 import strain_tail as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def stub_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("stub"))
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "stub_gpu.cpp"), os.path.join(ROOT, "oracle", "oracle_cpu.cpp"),
                           "-o", os.path.join(d, "libpantax_gpu.so"), "-lpthread"])
    return d


@pytest.fixture
def on_stub(stub_dir, monkeypatch, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the -m gpu tests run the driver on the real library")
    monkeypatch.setenv("LD_LIBRARY_PATH", stub_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    probe = str(tmp_path / "probe.gaf")
    open(probe, "w").close()
    r = subprocess.run([thd.BIN, "--filter-gaf", probe], capture_output=True, text=True)  # any run that creates a context
    assert r.returncode == 0, r.stderr  # ... succeeds: the stub is the library in use (the real one refuses without a device)


def test_driver_end_to_end_tables(on_stub, tmp_path):
    thd.test_host_driver_end_to_end_against_oracle.__wrapped__(tmp_path) if hasattr(thd.test_host_driver_end_to_end_against_oracle, "__wrapped__") \
        else thd.test_host_driver_end_to_end_against_oracle(tmp_path)


def test_driver_strain_only_resume_and_stdin(on_stub, tmp_path):
    thd.test_host_driver_strain_only_resume_and_stdin(tmp_path)


def test_driver_filter_gaf_file_mode(on_stub, tmp_path):
    tsc.test_zz_filter_gaf_file_mode(tmp_path)


def test_driver_then_strain_stage(on_stub, tmp_path):
    """pantax-gpu-profile --species --strain, then the strain stage on ITS files == the in-memory tail over the same numbers:
    the formats the driver writes are the formats the stage reads (strain_graphs for the GFA / lz4 / zst species included)."""
    tmp = str(tmp_path)
    ds = synth.Dataset(78, [3000, 2500, 1500, 1200], [5, 2, 3, 4])
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(6, 0, 40000)
    ranges = ds.ranges()
    db, gp, _lens = thd.make_db(tmp, ds, graphs, gaf)
    info = tsc.make_genomes_info(db, ranges, graphs)
    wd = os.path.join(tmp, "wd")
    os.makedirs(wd)
    r = subprocess.run([thd.BIN, "--db", db, "--gaf", gp, "--wd", wd, "--species", "--strain", "-a", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for s, (t, _a, _b) in enumerate(ranges):
        assert os.path.exists(os.path.join(wd, "strain_inputs", f"{t}.paths.tsv"))
        assert os.path.exists(os.path.join(wd, "strain_graphs", f"{t}.bin")) == (s != 0)
        if s != 0:  # the plain .bin the driver left is the species' Graph
            lens, names, paths = st.read_bin_graph(os.path.join(wd, "strain_graphs", f"{t}.bin"))
            assert lens.tolist() == graphs[s][0].tolist() and names == graphs[s][2]
            assert [p.tolist() for p in paths] == [np.asarray(p).tolist() for p in graphs[s][1]]
    args = st.ProfilingArgs()
    rows = st.run_strain_stage(db, wd, args, os.path.join(tmp, "ori.txt"))
    # the same tail without the driver's files: what the driver must have written, rebuilt from the restatement
    o = run_cpu_oracle(ranges, graphs, gaf)
    cov_of = {x[0]: float(x[2]) for x in st._tsv(os.path.join(wd, "species_abundance.txt"))}
    wd2 = os.path.join(tmp, "wd2")
    species = []
    for s, (t, _a, _b) in enumerate(ranges):
        lens, paths, names = graphs[s]
        depth, possible, om, ratio = tsc.write_strain_inputs(wd2, t, lens, paths, names, o, s)
        species.append((t, lens, paths, names, depth, possible, om, ratio))
        # ... and the driver's tables are those tables, value for value (frequencies_mean in the reference's trio order)
        for f in (f"{t}.nodes.tsv", f"{t}.paths.tsv"):
            a = st._tsv(os.path.join(wd, "strain_inputs", f))
            b = st._tsv(os.path.join(wd2, "strain_inputs", f))
            assert len(a) == len(b)
            for ra, rb in zip(a, b):
                assert len(ra) == len(rb)
                for x, y in zip(ra, rb):
                    assert x == y or float(x) == pytest.approx(float(y), rel=1e-12)
    want = tsc.in_memory_rows(args, species, cov_of, info, tmp)
    assert len(rows) >= 3 and [x[:3] for x in rows] == [x[:3] for x in want]
    for a, b in zip(rows, want):
        for x, y in zip(a[3:], b[3:]):
            assert (x == "" and y == "") or float(x) == pytest.approx(float(y), rel=1e-9, abs=1e-12)


def test_profile_command_runs_both_stages(on_stub, tmp_path):
    """`python -m pantax_b200 profile`: driver, then strain stage, in one command."""
    import sys

    tmp = str(tmp_path)
    ds = synth.Dataset(79, [2500, 1800], [4, 3])
    graphs = dataset_graphs(ds)
    gaf = ds.gaf(3, 0, 25000)
    db, gp, _lens = thd.make_db(tmp, ds, graphs, gaf)
    tsc.make_genomes_info(db, ds.ranges(), graphs)
    wd = os.path.join(tmp, "out")
    r = subprocess.run([sys.executable, "-m", "pantax_b200", "profile", "--db", db, "--gaf", gp, "--wd", wd, "-a", "0"], cwd=tmp, capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stderr
    for f in ("species_abundance.txt", "strain_abundance.txt"):
        assert os.path.getsize(os.path.join(wd, f)) > 40
    assert os.path.exists(os.path.join(tmp, "ori_strain_abundance.txt"))  # profile.rs:3217: the current directory
    # a second run finds both tables and leaves them alone (profile.rs:3414-3422)
    before = open(os.path.join(wd, "species_abundance.txt")).read()
    r = subprocess.run([thd.BIN, "--db", db, "--gaf", gp, "--wd", wd, "--species", "--strain"], capture_output=True, text=True)
    assert r.returncode == 0 and "both exist" in r.stderr and open(os.path.join(wd, "species_abundance.txt")).read() == before
