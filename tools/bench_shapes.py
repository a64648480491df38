"""Secondary measurements (not the bench line): throughput of the hot path on the other BASELINE.json shapes,
HBM-resident text, same step definition as bench.py.  Prints one JSON line per shape."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
from common import dataset_graphs
from pantax_b200 import api

cudart = C.CDLL("libcudart.so")


def run(name, ds, gseed, n_records, params, steps=5):
    graphs = dataset_graphs(ds)
    buf, n = ds.gaf_raw(gseed, 0, n_records, params)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    t0 = time.perf_counter()
    for s, g in enumerate(graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    t_graph = time.perf_counter() - t0
    ctx.reserve(n_records)
    bid, dptr = ctx.gaf_buffer_alloc(n)
    assert cudart.cudaMemcpy(C.c_void_p(dptr), buf, C.c_size_t(n), 1) == 0
    synth.lib().synth_free(buf)
    for _ in range(3):
        ctx.rewind(); ctx.ingest_gaf_device(bid, n); ctx.finalize()
    cudart.cudaDeviceSynchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.rewind(); ctx.ingest_gaf_device(bid, n); ctx.finalize()
    cudart.cudaDeviceSynchronize()
    dt = (time.perf_counter() - t0) / steps
    st = ctx.stats()
    print(json.dumps({"shape": name, "records": ctx.num_records, "gaf_bytes": n, "mean_line_bytes": n / ctx.num_records,
                      "ms_per_step": 1e3 * dt, "records_per_s": ctx.num_records / dt, "text_GBps": n / dt / 1e9,
                      "count_ms": st["count_ms"], "ingest_ms": st["ingest_ms"], "apply_ms": st["apply_ms"], "finalize_ms": st["finalize_ms"],
                      "nodes": st["nodes"], "paths": st["paths"], "unique_trios": st["unique_trios"], "graph_setup_s": t_graph,
                      "ids_unique": st["ids_unique"]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    # configs[2]: HiFi long reads (mean 15 kb, ~40-node walks) on a 100-species graph, 5 M nodes
    nodes = rng.multinomial(5_000_000, np.ones(100) / 100).tolist()
    haps = rng.integers(2, 9, size=100).tolist()
    if only in ("", "c3"):
      run("configs[2] HiFi 100 species 5M nodes 1M reads", synth.Dataset(20261017 + 3, nodes, haps, backbone_mean=400), 20261017 + 3, 1_000_000,
          synth.GafParams(long_reads=True, id_pair_suffix=False))
    # configs[3] scaled to one GPU: 1,000 species, 20 M nodes, 25 M short reads (one eighth of the 200 M)
    nodes = rng.multinomial(20_000_000, np.ones(1000) / 1000).tolist()
    haps = rng.integers(1, 9, size=1000).tolist()
    nrec = int(os.environ.get("C4_RECORDS", "25000000"))
    if only in ("", "c4"):
        run(f"configs[3] scaled: short reads 1000 species 20M nodes {nrec} reads", synth.Dataset(20261017 + 4, nodes, haps), 20261017 + 4, nrec, synth.GafParams())
