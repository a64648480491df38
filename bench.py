#!/usr/bin/env python
"""bench.py - GAF records/s to node coverage + strain statistics (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] - single-species synthetic pangenome,
1 M nodes, 50 strain paths, 10 M short-read GAF records (vg-giraffe dialect, 150 bp) per GPU.
Under torchrun (N>1) every rank owns its own 10 M-record read batch of the same stream (weak
scaling), the graph is replicated and ptx_finalize reduces over NCCL.

A step = one pass of the hot path over the batch: zeroed accumulators -> k_ingest (one pass over the
text: parse/classify/count -> record table + CSR walks; only the very first chunk of a ctx is preceded
by a record-count pass) -> k_apply (id set + node coverage + trio sums) -> finalize (covered bases, per-path sums,
per-hap unique-trio counts).  The graph upload + unique-trio table build is database setup
(SURVEY.md section 8d), timed separately and reported in config.

  value  : GAF text already resident in HBM when the timed region starts (ptx_ingest_gaf_device)
  e2e    : same metric through the C ABI with HOST buffers - ptx_ingest_gaf from pinned memory
           (H2D inside the timed region) + D2H of the per-node / per-path / per-hap results
  roofline: dominant kernel k_ingest, algorithmic bytes / its CUDA-event duration
  cpu_baseline: the C++ oracle port of the reference (all host threads) on a bounded sample

--impl reference times that CPU port alone (the reference itself is Rust and cannot be built
in this image - see DESIGN.md); it never touches the GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "GAF records/s to node coverage+strain stats"
SEED = 20261017 + 2  # seed = 20261017 + config#


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=10_000_000, help="GAF records per GPU")
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--haps", type=int, default=50)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="records in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_dataset(args):
    import synth

    ds = synth.Dataset(SEED, [args.nodes], [args.haps])
    return ds


def workload_name(args, n_gpus):
    return (f"BASELINE configs[1]: single-species synthetic graph, {args.nodes} nodes, {args.haps} strain paths, "
            f"{args.records} short-read GAF records per GPU x {n_gpus} GPU")


# --------------------------------------------------------------------------------------
# reference arm: the CPU port of the reference's path, all host threads
# --------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    import synth
    from common import dataset_graphs
    from oracle import cpu as ocpu

    ds = make_dataset(args)
    graphs = dataset_graphs(ds)
    o = ocpu.CpuOracle(0)
    o.set_ranges(ds.ranges())
    o.set_graph(0, graphs[0][0], graphs[0][1])
    t_prep = o.prepare_graphs()
    sample = min(args.records, args.cpu_sample)
    buf, nbytes = ds.gaf_raw(SEED, 0, sample)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.run(buf.value, nbytes)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    nrec = o.n_records
    synth.lib().synth_free(buf)
    total = sum(times)
    value = nrec * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample": f"each step = first {nrec} records of rank 0's batch",
                   "graph_setup_s": t_prep},
        "cpu_baseline": {"value": value, "unit": "records/s", "cores": o.threads, "kind": "port",
                         "sample": f"first {nrec} records of the workload per step; C++ port of the reference "
                                   f"(oracle/oracle_cpu.cpp), {o.threads} threads; the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import synth
    from common import dataset_graphs
    from pantax_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (the library has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic inputs: graph replicated, rank r owns records [r*R, (r+1)*R)
    ds = make_dataset(args)
    graphs = dataset_graphs(ds)
    R = args.records
    buf, nbytes = ds.gaf_raw(SEED, rank * R, (rank + 1) * R)
    pinned = api.PinnedBuffer(nbytes)
    C.memmove(pinned.ptr, buf, nbytes)
    synth.lib().synth_free(buf)
    text = pinned.array
    walk_nodes = int(np.count_nonzero(text == ord(">")) + np.count_nonzero(text == ord("<")))

    ctx = api.PantaxGpu(local_rank)
    ctx.set_ranges(ds.ranges())
    t0 = time.perf_counter()
    ctx.upload_graph(0, graphs[0][0], graphs[0][1])
    ctx.commit_graphs()
    t_graph = time.perf_counter() - t0
    ctx.reserve(R)  # before comm_init: sizes the peer-memory id boxes
    if world > 1:
        uid = [api.PantaxGpu.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    cudart = C.CDLL("libcudart.so")
    bid, dptr = ctx.gaf_buffer_alloc(nbytes)
    assert cudart.cudaMemcpy(C.c_void_p(dptr), C.c_void_p(pinned.ptr), C.c_size_t(nbytes), 1) == 0

    def step_resident():
        ctx.rewind()
        ctx.ingest_gaf_device(bid, nbytes)
        ctx.finalize()

    # ---- value: text resident in HBM
    for _ in range(max(args.warmup, 3)):
        step_resident()
    n_rec = ctx.num_records
    _, _, launches0 = ctx.timing()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    kern_ms = []
    for _ in range(args.steps):
        step_resident()
        st = ctx.stats()  # CUDA-event times of this step's kernels (events on the library's stream)
        kern_ms.append((st["count_ms"], st["ingest_ms"], st["finalize_ms"], st["ingest_launches"], st.get("apply_ms", 0.0)))
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    _, _, launches1 = ctx.timing()
    total_records = sum_over_ranks(float(n_rec))
    value = total_records * args.steps / dt
    ms_per_step = 1e3 * dt / args.steps
    gpu_launches = int(launches1 - launches0)

    # ---- roofline of the dominant kernel (k_ingest): algorithmic bytes / CUDA-event duration
    count_ms = float(np.mean([k[0] for k in kern_ms]))
    ingest_ms = float(np.mean([k[1] for k in kern_ms]))
    final_ms = float(np.mean([k[2] for k in kern_ms]))
    apply_ms = float(np.mean([k[4] for k in kern_ms]))
    n_launch = max(1, int(kern_ms[-1][3]))
    peak, peak_src = peaks()

    # ---- cpu baseline + workload statistics (trio probe hit rate) on a bounded sample, rank 0, N=1 only
    cpu = None
    p_hit = 0.0
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        from oracle import cpu as ocpu

        o = ocpu.CpuOracle(0)
        o.set_ranges(ds.ranges())
        o.set_graph(0, graphs[0][0], graphs[0][1])
        o.prepare_graphs()
        sample = min(R, args.cpu_sample)
        sbuf, sbytes = ds.gaf_raw(SEED, 0, sample)
        o.run(sbuf.value, sbytes)  # warm
        t0 = time.perf_counter()
        o.run(sbuf.value, sbytes)
        cdt = time.perf_counter() - t0
        ws = o.workload_stats()
        p_hit = ws["trio_hits"] / max(1, ws["trio_windows"])
        # parity of the timed configuration itself, on the sample prefix (bit-exact integers)
        cpu = {"value": o.n_records / cdt, "unit": "records/s", "cores": o.threads, "kind": "port",
               "sample": f"first {o.n_records} of {n_rec} records; oracle/oracle_cpu.cpp (C++ port of the reference, "
                         f"{o.threads} threads); stages s: parse+classify+counts {o.times()[0]:.3f}, id grouping {o.times()[1]:.3f}, "
                         f"coverage {o.times()[2]:.3f}, stats {o.times()[3]:.3f}"}
        synth.lib().synth_free(sbuf)
        del o

    L = nbytes / n_rec
    W = walk_nodes / n_rec
    # B_rec = L + 12 W + 8 max(W-2,0) p_hit.  The path runs as two kernels: k_ingest reads the text
    # (L bytes per record), k_apply does the node/trio accumulation (the 12 W + 8 (W-2) p_hit part).
    b_apply = 12.0 * W + 8.0 * max(W - 2.0, 0.0) * p_hit
    b_rec = L + b_apply
    alg_bytes = n_rec * L / n_launch          # dominant kernel: k_ingest
    achieved = alg_bytes / (ingest_ms / n_launch * 1e-3) / 1e9
    achieved_apply = n_rec * b_apply / max(apply_ms, 1e-9) / 1e6
    achieved_path = n_rec * b_rec / max(ingest_ms + apply_ms, 1e-9) / 1e6
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_ingest_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("records") == n_rec:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "k_ingest (GAF parse -> record table + CSR walks)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_record": L, "mean_line_bytes": L, "mean_walk_nodes": W, "trio_hit_rate": p_hit,
                "kernel_ms": ingest_ms / n_launch, "count_ms": count_ms, "finalize_ms": final_ms,
                "k_apply": {"kernel_ms": apply_ms / n_launch, "algorithmic_bytes_per_record": b_apply, "achieved": achieved_apply,
                            "frac": achieved_apply / peak},
                "whole_path": {"kernels": "k_ingest + k_apply", "algorithmic_bytes_per_record": b_rec, "achieved": achieved_path,
                               "frac": achieved_path / peak}}

    # ---- e2e: host buffers through the C ABI, H2D inside the timed region, results read back
    e2e = None
    if not args.no_e2e:
        ctx.reset()  # releases the resident buffer

        # result arrays in pinned host memory, allocated once by the caller like the input buffer
        n_nodes0, n_trios0 = ctx.n_nodes(0), ctx.n_trios(0)
        res_pin = api.PinnedBuffer(8 * (2 * max(n_nodes0, 1) + max(n_trios0, 1)))
        o_bases = res_pin.view(np.int64, max(n_nodes0, 1))
        o_cov = res_pin.view(np.uint64, max(n_nodes0, 1), 8 * max(n_nodes0, 1))
        o_trio = res_pin.view(np.int64, max(n_trios0, 1), 16 * max(n_nodes0, 1))

        def step_e2e():
            ctx.reset()
            ctx.ingest_gaf(pinned, is_last=True)
            ctx.finalize()
            out = [ctx.species_counts(), ctx.node_bases(0, out=o_bases), ctx.node_cov(0, out=o_cov), ctx.trio_bases(0, out=o_trio)]
            out += list(ctx.path_sums(0)) + list(ctx.hap_trio_counts(0))
            return out

        for _ in range(2):
            out = step_e2e()
        d2h = int(sum(a.nbytes for a in out))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = step_e2e()
        barrier()
        edt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_records * args.steps / edt, "unit": "records/s", "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * edt / args.steps}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "records_per_gpu": int(n_rec), "gaf_bytes_per_gpu": int(nbytes),
                       "l2": f"input text {nbytes / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "graph_setup_s": t_graph, "unique_trios": ctx.n_trios(0), "timing": "wall clock over K steps between "
                       "barrier+synchronize, max over ranks; kernel times from CUDA events on the library stream"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
