"""Times k_ingest<CLASSIFY> (no coverage) and k_ingest<COVER> (replay) separately on config 2 (experiment)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
from common import dataset_graphs
from pantax_b200 import api
ds = synth.Dataset(20261019, [1000000], [50]); graphs = dataset_graphs(ds)
buf, n = ds.gaf_raw(20261019, 0, 10_000_000)
cudart = C.CDLL("libcudart.so")
for rep in range(3):
    ctx = api.PantaxGpu(0); ctx.set_ranges(ds.ranges()); ctx.reserve(10_000_000)
    bid, dptr = ctx.gaf_buffer_alloc(n)
    assert cudart.cudaMemcpy(C.c_void_p(dptr), buf, C.c_size_t(n), 1) == 0
    ctx.ingest_gaf_device(bid, n); ctx.finalize()
    s1 = ctx.stats()
    ctx.upload_graph(0, graphs[0][0], graphs[0][1]); ctx.commit_graphs(); ctx.finalize()
    s2 = ctx.stats()
    print("classify-only ingest_ms", s1["ingest_ms"], "| cover-only (in finalize) ms", s2["finalize_ms"] - s1["finalize_ms"], flush=True)
    ctx.close()
