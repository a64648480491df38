"""Multi-GPU parity as a `-m gpu` test: launches tools/multigpu_check.py under torchrun on min(#GPUs, 4) ranks of this box.
Every rank ingests its read-batch shard, ptx_finalize reduces over NCCL, and every rank's result must equal the C++
oracle's on the WHOLE input bit for bit - with unique ids, ids duplicated across shards (mixed-species groups),
the outbox-overflow restart, each over NCCL boxes and over peer-memory boxes.  Skipped on a box with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_sharded_ranks_reduce_to_the_single_process_oracle():
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs on one box (found {n})")
    ranks = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", "29653", os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert r.stdout.count("bit-exact on every rank") == 6, r.stdout[-3000:]


@pytest.mark.gpu
def test_two_contexts_on_two_devices_in_one_process():
    """The reference CLI is ONE process: it must be able to drive several GPUs (one ctx each) from it.  The opt-in
    shared-memory size of the ingest kernels is a per-device function attribute."""
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs on one box (found {n})")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from common import NASTY, dataset_graphs, run_cpu_oracle, synth
    from gpu_common import assert_gpu_matches_oracle
    from pantax_b200 import api

    ds = synth.Dataset(92, [30000, 8000], [6, 2])
    gaf = ds.gaf(5, 0, 60000, NASTY)
    graphs = dataset_graphs(ds)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    for dev in (0, 1):
        ctx = api.PantaxGpu(dev)
        ctx.set_ranges(ds.ranges())
        for s, g in enumerate(graphs):
            ctx.upload_graph(s, g[0], g[1])
        ctx.commit_graphs()
        ctx.ingest_gaf(gaf, is_last=True)
        ctx.finalize()
        assert_gpu_matches_oracle(ctx, o, graphs)
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("reserve", [0, 60000])
def test_create_multi_drives_all_gpus_from_one_process(reserve):
    """ptx_create_multi / ptx_finalize_multi: N contexts in ONE process (ncclCommInitAll, peer-access id boxes when a
    size hint is given), each fed its read-batch shard; the reduced result on every context is the oracle's."""
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs on one box (found {n})")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import NASTY_DUP, dataset_graphs, run_cpu_oracle, synth
    from gpu_common import assert_gpu_matches_oracle
    from pantax_b200 import api
    from pantax_b200.shard import shard_bounds_bytes

    P = min(n, 4)
    ds = synth.Dataset(93, [40000, 10000, 3000], [8, 3, 1])
    gaf = ds.gaf(6, 0, 120000, NASTY_DUP)
    graphs = dataset_graphs(ds)
    o = run_cpu_oracle(ds.ranges(), graphs, gaf)
    ctxs = api.PantaxGpu.create_multi(list(range(P)), reserve)
    try:
        for ctx in ctxs:
            ctx.set_ranges(ds.ranges())
            for s, g in enumerate(graphs):
                ctx.upload_graph(s, g[0], g[1])
            ctx.commit_graphs()
        for ctx, (lo, hi) in zip(ctxs, shard_bounds_bytes(gaf, P)):
            ctx.ingest_gaf(gaf[lo:hi], is_last=True)
        api.PantaxGpu.finalize_multi(ctxs)
        assert sum(c.num_records for c in ctxs) == o.n_records
        assert bool(ctxs[0].stats().get("p2p_boxes")) == bool(reserve)
        for ctx in ctxs:
            try:
                api.PantaxGpu.num_records = property(lambda self: o.n_records)
                assert_gpu_matches_oracle(ctx, o, graphs, check_labels=False)
            finally:
                api.PantaxGpu.num_records = property(lambda self: self._L.ptx_num_records(self._h))
    finally:
        for ctx in ctxs:
            ctx.close()


@pytest.mark.gpu
def test_equal_length_takes_the_first_1000_rows_of_the_whole_gaf_not_of_rank_0():
    """profile.rs:311-322 looks at the first 1000 non-U rows of the GAF.  With the reads sharded by batch, rank 0 may hold
    fewer than 1000 of them: ptx_finalize then gathers the rest from the ranks behind it, in rank order.  Here every read
    of the first two shards is 150 bp and one read in the third shard (row ~700) is not."""
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs on one box (found {n})")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import dataset_graphs, synth
    from oracle import pantax_oracle as opy
    from pantax_b200 import api
    from pantax_b200.shard import shard_bounds_bytes

    P = min(n, 4)
    ds = synth.Dataset(94, [20000, 5000], [4, 2])
    graphs = dataset_graphs(ds)
    base = ds.gaf(7, 0, 900, synth.GafParams())
    for odd_at in (700, 850):  # inside a later shard, within the first 1000 rows
        lines = base.split(b"\n")
        f = lines[odd_at].split(b"\t")
        f[1] = b"149"
        lines[odd_at] = b"\t".join(f)
        gaf = b"\n".join(lines)
        rows = opy.rcls_profile(gaf, ds.ranges())
        want = opy.equal_length_test(rows)
        assert want[0] is False  # (the odd row is classified; otherwise the case tests nothing)
        for data, expect in ((gaf, want), (base, opy.equal_length_test(opy.rcls_profile(base, ds.ranges())))):
            ctxs = api.PantaxGpu.create_multi(list(range(P)), 0)
            try:
                for ctx in ctxs:
                    ctx.set_ranges(ds.ranges())
                    for s, g in enumerate(graphs):
                        ctx.upload_graph(s, g[0], g[1])
                    ctx.commit_graphs()
                bounds = shard_bounds_bytes(data, P)
                assert bounds[0][1] < bounds[-1][1]  # rank 0 holds only a part of the 900 rows
                for ctx, (lo, hi) in zip(ctxs, bounds):
                    ctx.ingest_gaf(data[lo:hi], is_last=True)
                api.PantaxGpu.finalize_multi(ctxs)
                for ctx in ctxs:
                    eq, rl = ctx.equal_length()
                    assert (eq, rl if eq else None) == expect
            finally:
                for ctx in ctxs:
                    ctx.close()
