"""Scatter-variant bake-off of the node-coverage pass (north_star stage 2; replaces the read loop of profile.rs:798-883).

    python tools/bench_scatter.py <shape> <variant> [steps]      shape: c1 | c2 | n50m      variant: 0 | 1 | 2 | 3

One (shape, variant) per process (PTX_SCATTER is read at ptx_create).  Prints one JSON line with the CUDA-event times of
the step's kernels.  Under `ncu --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum` the same command gives the per-kernel counters of profiles/r2_scatter_bakeoff.md.
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    shape, var = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    os.environ["PTX_SCATTER"] = str(var)
    import numpy as np
    import synth
    from common import dataset_graphs
    from pantax_b200 import api
    import bench

    if shape == "c1":
        ds, params, n_rec = synth.Dataset(20261019, [1_000_000], [50]), synth.GafParams(), 10_000_000
    elif shape == "c2":
        ds, params, n_rec = synth.Dataset(20261020, [50_000] * 100, [5] * 100, backbone_mean=300.0), synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.1), 1_000_000
    elif shape == "n50m":
        ds, params, n_rec = synth.Dataset(20261022, [50_000] * 1000, [5] * 1000), synth.GafParams(), 20_000_000
    elif shape == "hot":  # few nodes, many reads: the shape the shared-memory table is for
        ds, params, n_rec = synth.Dataset(20261023, [4_000], [4]), synth.GafParams(), 5_000_000
    else:
        raise SystemExit("unknown shape")
    graphs = dataset_graphs(ds)
    ctx = api.PantaxGpu(0)
    ctx.set_ranges(ds.ranges())
    t0 = time.perf_counter()
    for s, g in enumerate(graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    t_graph = time.perf_counter() - t0
    ctx.reserve(n_rec)
    cudart = C.CDLL("libcudart.so")
    wl = bench.Workload(shape, shape, 7, [1], [1], params)
    wl._ds = ds
    bufs, nbytes, walk = bench.load_resident(ctx, wl, 0, n_rec, cudart)
    for _ in range(3):
        bench.step_resident(ctx, bufs)
    t0 = time.perf_counter()
    k = []
    for _ in range(steps):
        bench.step_resident(ctx, bufs)
        st = ctx.stats()
        k.append((st["ingest_ms"], st["apply_ms"], st["finalize_ms"]))
    cudart.cudaDeviceSynchronize()
    dt = time.perf_counter() - t0
    k = np.array(k)
    bases_sum = int(sum(int(ctx.node_bases(s).sum()) for s in range(min(ds.n_species, 3))))
    print(json.dumps({"shape": shape, "variant": var, "records": int(ctx.num_records), "nodes": int(ds.total_nodes), "walk_nodes": walk, "steps": steps,
                      "ms_per_step": 1e3 * dt / steps, "ingest_ms": float(k[:, 0].mean()), "apply_ms": float(k[:, 1].mean()),
                      "finalize_ms": float(k[:, 2].mean()), "graph_setup_s": t_graph, "bases_checksum_first3": bases_sum}), flush=True)


if __name__ == "__main__":
    main()
