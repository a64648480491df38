"""Where the end-to-end step goes: raw pinned->device copy rate of the same bytes vs the C-ABI e2e step."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
from common import dataset_graphs
from pantax_b200 import api

cudart = C.CDLL("libcudart.so")
n_records = int(os.environ.get("RECORDS", "10000000"))
ds = synth.Dataset(20261017 + 2, [1_000_000], [50])
graphs = dataset_graphs(ds)
buf, n = ds.gaf_raw(20261017 + 2, 0, n_records, synth.GafParams())
ctx = api.PantaxGpu(0)
ctx.set_ranges(ds.ranges())
for s, g in enumerate(graphs): ctx.upload_graph(s, g[0], g[1])
ctx.commit_graphs(); ctx.reserve(n_records)
pin = api.PinnedBuffer(n)
C.memmove(pin.ptr, buf, n)
dptr = C.c_void_p()
assert cudart.cudaMalloc(C.byref(dptr), C.c_size_t(n)) == 0
for piece in (n, 64 << 20):
    for rep in range(3):
        cudart.cudaDeviceSynchronize(); t0 = time.perf_counter()
        off = 0
        while off < n:
            k = min(piece, n - off)
            cudart.cudaMemcpyAsync(C.c_void_p(dptr.value + off), C.c_void_p(pin.ptr + off), C.c_size_t(k), 1, None)
            off += k
        cudart.cudaDeviceSynchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"raw_h2d_piece": piece, "ms": 1e3 * dt, "GBps": n / dt / 1e9}), flush=True)
for rep in range(4):
    cudart.cudaDeviceSynchronize(); t0 = time.perf_counter()
    ctx.reset(); t1 = time.perf_counter()
    ctx.ingest_gaf(pin, is_last=True); t2 = time.perf_counter()
    ctx.finalize(); t3 = time.perf_counter()
    b = ctx.node_bases(0); c = ctx.node_cov(0); t4 = time.perf_counter()
    print(json.dumps({"e2e_ms": 1e3 * (t4 - t0), "reset": 1e3 * (t1 - t0), "ingest": 1e3 * (t2 - t1), "finalize": 1e3 * (t3 - t2), "d2h": 1e3 * (t4 - t3)}), flush=True)
