"""Helpers for the `-m gpu` parity tests: run the CUDA path through the C ABI and compare
every integer output with the C++ oracle bit for bit."""
from __future__ import annotations

import numpy as np

from common import run_cpu_oracle
from pantax_b200 import api


def run_gpu(ranges, graphs, gaf: bytes, split=None, flow="fused", reserve=None):
    """flow='fused': graphs first, coverage fused into the ingest;
    flow='late': the reference's order - classify, then upload graphs, then coverage replay."""
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ranges)
    if reserve:
        ctx.reserve(reserve)

    def upload():
        any_graph = False
        for s, g in enumerate(graphs):
            if g is not None:
                ctx.upload_graph(s, g[0], g[1])
                any_graph = True
        if any_graph:
            ctx.commit_graphs()

    if flow == "fused":
        upload()
    if split:
        cuts = [0] + sorted(split) + [len(gaf)]
        for i in range(len(cuts) - 1):
            ctx.ingest_gaf(gaf[cuts[i]:cuts[i + 1]], is_last=(i == len(cuts) - 2))
    else:
        ctx.ingest_gaf(gaf, is_last=True)
    ctx.finalize()
    if flow == "late":
        upload()
        ctx.finalize()
    return ctx


def assert_gpu_matches_oracle(ctx, o, graphs, check_labels=True):
    assert ctx.num_records == o.n_records
    if check_labels:
        np.testing.assert_array_equal(ctx.read_labels(), o.labels())
    np.testing.assert_array_equal(ctx.species_counts(), o.species_counts())
    assert ctx.ids_unique == o.ids_unique
    for s, g in enumerate(graphs):
        if g is None:
            continue
        if o.species_error(s):
            try:
                ctx.node_bases(s)
                raise AssertionError("expected PTX_E_START_GT_LEN")
            except api.PantaxGpuError as e:
                assert e.name == "PTX_E_START_GT_LEN"
            continue
        keys, tlen, owner = o.trio_table(s)
        gk, gl, go = ctx.trio_table(s)
        np.testing.assert_array_equal(gk, keys)
        np.testing.assert_array_equal(gl, tlen)
        np.testing.assert_array_equal(go, owner)
        bases = o.node_bases(s)
        np.testing.assert_array_equal(ctx.node_bases(s), bases)
        np.testing.assert_array_equal(ctx.node_cov(s), o.node_cov(s))
        tb = o.trio_bases(s)
        np.testing.assert_array_equal(ctx.trio_bases(s), tb)
        sc, sl = o.path_sums(s)
        gsc, gsl = ctx.path_sums(s)
        np.testing.assert_array_equal(gsc, sc)
        np.testing.assert_array_equal(gsl, sl)
        U, nz = o.hap_trio_counts(s)
        gU, gnz = ctx.hap_trio_counts(s)
        np.testing.assert_array_equal(gU, U)
        np.testing.assert_array_equal(gnz, nz)
        # the single IEEE divisions (profile.rs:988, :1014) must be bit-identical to the host's
        np.testing.assert_array_equal(ctx.node_depth(s), bases.astype(np.float64) / np.asarray(g[0], dtype=np.float64))
        if len(tb):
            np.testing.assert_array_equal(ctx.trio_depth(s), tb.astype(np.float64) / tlen.astype(np.float64))


def gpu_vs_oracle(ranges, graphs, gaf, **kw):
    o = run_cpu_oracle(ranges, graphs, gaf)
    ctx = run_gpu(ranges, graphs, gaf, **kw)
    assert_gpu_matches_oracle(ctx, o, graphs)
    return ctx, o
