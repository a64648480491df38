import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import dataset_graphs, synth
from oracle import pantax_oracle as opy
from pantax_b200 import api
from pantax_b200.shard import shard_bounds_bytes
P = 2
ds = synth.Dataset(94, [20000, 5000], [4, 2])
graphs = dataset_graphs(ds)
base = ds.gaf(7, 0, 900, synth.GafParams())
lines = base.split(b"\n"); f = lines[700].split(b"\t"); f[1] = b"149"; lines[700] = b"\t".join(f); gaf = b"\n".join(lines)
for data in (gaf, base):
    rows = opy.rcls_profile(data, ds.ranges())
    print("expect", opy.equal_length_test(rows))
    ctxs = api.PantaxGpu.create_multi(list(range(P)), 0)
    for ctx in ctxs:
        ctx.set_ranges(ds.ranges())
        for s, g in enumerate(graphs): ctx.upload_graph(s, g[0], g[1])
        ctx.commit_graphs()
    bounds = shard_bounds_bytes(data, P)
    print(bounds, len(data))
    for ctx, (lo, hi) in zip(ctxs, bounds): ctx.ingest_gaf(data[lo:hi], is_last=True)
    api.PantaxGpu.finalize_multi(ctxs)
    for ctx in ctxs: print(ctx.equal_length(), ctx.num_records, ctx.species_counts()[:2])
    for ctx in ctxs: ctx.close()
