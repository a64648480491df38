"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.  Every rank runs the ORACLE on its
read-batch shard; the int64 accumulators are all-reduced (sum) and the covered-base bitmaps OR-ed, exactly
what ptx_finalize does over NCCL - the result must equal the single-shard oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import NASTY, dataset_graphs, run_cpu_oracle, synth
from pantax_b200.shard import shard_bounds_bytes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = synth.Dataset(77, [6000, 2000], [5, 2])
    gaf = ds.gaf(3, 0, 6000, NASTY)  # ids unique: the id-group rule needs the cross-rank exchange
    graphs = dataset_graphs(ds)
    lo, hi = shard_bounds_bytes(gaf, world)[rank]
    o = run_cpu_oracle(ds.ranges(), graphs, gaf[lo:hi], threads=1)
    counts = torch.from_numpy(o.species_counts().copy())
    dist.all_reduce(counts)
    res = {"counts": counts.numpy(), "n": torch.tensor([o.n_records])}
    dist.all_reduce(res["n"])
    for s in range(2):
        b = torch.from_numpy(o.node_bases(s).copy())
        t = torch.from_numpy(o.trio_bases(s).copy())
        dist.all_reduce(b)
        dist.all_reduce(t)
        res[f"bases{s}"] = b.numpy()
        res[f"trio{s}"] = t.numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), **{k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in res.items()})
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_oracle_plus_allreduce_equals_single_shard(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    red = np.load(os.path.join(tmp_path, "reduced.npz"))
    ds = synth.Dataset(77, [6000, 2000], [5, 2])
    gaf = ds.gaf(3, 0, 6000, NASTY)
    o = run_cpu_oracle(ds.ranges(), dataset_graphs(ds), gaf, threads=2)
    assert int(red["n"][0]) == o.n_records
    np.testing.assert_array_equal(red["counts"], o.species_counts())
    for s in range(2):
        np.testing.assert_array_equal(red[f"bases{s}"], o.node_bases(s))
        np.testing.assert_array_equal(red[f"trio{s}"], o.trio_bases(s))


def test_shard_bounds_cover_every_line_exactly_once():
    ds = synth.Dataset(5, [3000], [4])
    gaf = ds.gaf(1, 0, 1000, NASTY)
    for world in (1, 2, 3, 8, 64, 5000):
        b = shard_bounds_bytes(gaf, world)
        assert b[0][0] == 0 and b[-1][1] == len(gaf)
        for (a0, a1), (b0, _b1) in zip(b, b[1:]):
            assert a1 == b0 and a0 <= a1
        for lo, hi in b:
            assert lo == 0 or lo == len(gaf) or gaf[lo - 1:lo] == b"\n"
        assert b"".join(gaf[lo:hi] for lo, hi in b) == gaf


def _id_hash(read_id: bytes) -> int:
    import zlib
    return (zlib.crc32(read_id) << 32) | zlib.adler32(read_id)   # any deterministic hash: the protocol only needs agreement


def _worker_idgroups(rank, world, port, out_dir):
    """The id-group protocol of ptx_finalize with plain Python objects over gloo: own ids are
    inserted locally, foreign ones travel to their owner, the owner's MIXED ids come back, every rank then covers
    its own reads with the keep mask and the accumulators are summed."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import NASTY_DUP, opy, py_graph
    from id_owner_model import IdOwnerSet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = synth.Dataset(78, [3000, 1000], [4, 2])
    gaf = ds.gaf(5, 0, 3000, NASTY_DUP)   # ids repeat, also across shards and across species
    graphs = dataset_graphs(ds)
    ranges = ds.ranges()
    name_to_idx = {r[0]: i for i, r in enumerate(ranges)}
    lo, hi = shard_bounds_bytes(gaf, world)[rank]
    rows = [r for r in opy.rcls_profile(gaf[lo:hi], ranges) if r.species != opy.UNCLASSIFIED]
    ids = IdOwnerSet(rank, world)
    elig = []
    for r in rows:
        e = None not in (r.path, r.read_path_len, r.read_start, r.read_end)   # profile.rs:380-399
        elig.append(e)
        ids.add_row(_id_hash(r.read_id), name_to_idx[r.species], e)
    boxes = [None] * world
    dist.all_gather_object(boxes, ids.outbox)             # stands for the grouped send/recv (or the peer-memory stores)
    for q in range(world):
        if q != rank:
            ids.merge_inbox(boxes[q][rank])
    flags = torch.tensor([int(ids.repeat), int(bool(ids.mixed_ids()))])
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    mixed_all = [None] * world
    dist.all_gather_object(mixed_all, ids.mixed_ids())    # return_mixed_ids
    mixed = set(h for part in mixed_all for h in part)
    res = {"flags": flags.numpy()}
    start_of = {name: s for name, s, _e in ranges}
    for s, (name, _a, _b) in enumerate(ranges):
        g = py_graph(*graphs[s])
        trio_map, trio_len, _owner = opy.trio_nodes_info(g)
        recs = [opy.Record(r.read_id, r.path, r.read_path_len, r.read_start, r.read_end, r.species)
                for r, e in zip(rows, elig) if e and r.species == name and _id_hash(r.read_id) not in mixed]
        bases, trio_bases, _cov, _na, _ta = opy.get_node_abundances(g.nodes_len, trio_map, trio_len, start_of[name], recs)
        b = torch.tensor(bases, dtype=torch.int64)
        t = torch.tensor(trio_bases if len(trio_bases) else [0], dtype=torch.int64)
        dist.all_reduce(b)
        dist.all_reduce(t)
        res[f"bases{s}"] = b.numpy()
        res[f"trio{s}"] = t.numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "idgroups.npz"), **res)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_id_group_ownership_protocol_equals_single_process(world, tmp_path):
    from common import NASTY_DUP
    mp.spawn(_worker_idgroups, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    red = np.load(os.path.join(tmp_path, "idgroups.npz"))
    ds = synth.Dataset(78, [3000, 1000], [4, 2])
    gaf = ds.gaf(5, 0, 3000, NASTY_DUP)
    o = run_cpu_oracle(ds.ranges(), dataset_graphs(ds), gaf, threads=2)
    assert bool(red["flags"][0]) == (not o.ids_unique)
    assert bool(red["flags"][1]) == (o.mixed_dropped > 0)
    assert o.mixed_dropped > 0   # the case is only interesting if groups are really dropped
    for s in range(2):
        np.testing.assert_array_equal(red[f"bases{s}"], o.node_bases(s))
        tb = o.trio_bases(s)
        np.testing.assert_array_equal(red[f"trio{s}"][: len(tb)], tb)
