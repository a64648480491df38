#!/usr/bin/env python
"""bench.py - GAF records/s to node coverage + strain statistics (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (config.workload): BASELINE.json configs[1] - single-species synthetic pangenome,
1 M nodes, 50 strain paths, 10 M short-read GAF records (vg-giraffe dialect, 150 bp) per GPU.
Under torchrun (N>1) every rank owns its own 10 M-record read batch of the same stream (weak
scaling), the graph is replicated and ptx_finalize reduces over NCCL.

A step = one pass of the hot path over the batch: zeroed accumulators -> k_ingest_s (one pass over the
text: structural index, parse/classify/count -> record table + CSR walks) -> k_apply (id set + node
coverage + trio sums) -> finalize (covered bases, per-path sums, per-hap unique-trio counts).  The
graph upload + unique-trio table build is database setup, timed separately.

  value   : GAF text already resident in HBM when the timed region starts (ptx_ingest_gaf_device)
  e2e     : same metric through the C ABI with HOST buffers - ptx_ingest_gaf from pinned memory
            (H2D inside the timed region) + D2H of the per-node / per-path / per-hap results
  roofline: dominant kernel k_ingest_s, algorithmic bytes / its CUDA-event duration
  cpu_baseline: the C++ oracle port of the reference (all host threads) on the whole workload of rank 0
  parity  : checked INSIDE this run (outside the timed regions): the timed context's integer outputs against the
            C++ oracle on the same records; resident single-chunk result == chunked host-path result; at N>1 every
            rank's reduced vectors == rank 0's == a 1-GPU streamed run over all N batches
  secondary: one more BASELINE config per run, same step definition, with its own parity check -
            N=1: configs[2] (HiFi long reads, 100 species, 5 M nodes, 1 M records);
            N>1: configs[3] (1,000 species, 20 M nodes, 200 M records sharded over the ranks)
  tertiary: at 8 GPUs (or --config4): configs[4] (50 M nodes, 5,000 paths, 1 B records over the ranks), parity on a 1/512 prefix

--impl reference times the CPU port alone (the reference itself is Rust and cannot be built in this
image - see DESIGN.md); it never touches the GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=10_000_000, help="GAF records per GPU (configs[1])")
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--haps", type=int, default=50)
    ap.add_argument("--cpu-sample", type=int, default=10_000_000, help="records in the cpu_baseline / reference-arm sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary-records", type=int, default=0, help="total records of the secondary config (0 = the BASELINE size)")
    ap.add_argument("--config4", action="store_true", help="also run BASELINE configs[4] (50 M nodes, 1 B records over the ranks); always on at 8 GPUs")
    ap.add_argument("--config4-records", type=int, default=0, help="total records of configs[4] (0 = 1 B)")
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
    """SM clock + clock-event reasons DURING the timed region: NVML polled every 2 ms from a thread (a timed region
    of 20 steps lasts ~40 ms - nvidia-smi's 100 ms loop never fired inside it), nvidia-smi as the fallback."""

    NAMES = [("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"), ("sw_thermal_slowdown", "SwThermalSlowdown"),
             ("sw_power_cap", "SwPowerCap")]

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.mx = [], set(), None
        self.stop_flag = False
        self.thread = None
        self.nvml = None
        self.how = "none"

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(self.get_reasons(self.h))
                for name, key in self.NAMES:
                    bit = getattr(n, "nvmlClocksEventReason" + key, None) or getattr(n, "nvmlClocksThrottleReason" + key, 0)
                    if r & int(bit):
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as n

            n.nvmlInit()
            phys = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            self.h = n.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
            self.get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml = n
            self.how = "nvml, 2 ms poll"
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if not self.sm:  # NVML unavailable: one nvidia-smi sample right behind the region
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.sm.append(float(out[0]))
                self.mx = float(out[1])
                for (name, _k), v in zip(self.NAMES, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
                self.how = "nvidia-smi, one sample behind the region"
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


# --------------------------------------------------------------------------------------
# workloads (BASELINE.json configs; SURVEY.md section 8d shapes)
# --------------------------------------------------------------------------------------
class Workload:
    def __init__(self, key, name, seed, nodes, haps, params, backbone_mean=64.0):
        self.key, self.name, self.seed, self.nodes, self.haps, self.params, self.backbone_mean = key, name, seed, nodes, haps, params, backbone_mean
        self._ds = None
        self._graphs = None

    @property
    def ds(self):
        import synth

        if self._ds is None:
            self._ds = synth.Dataset(self.seed, self.nodes, self.haps, backbone_mean=self.backbone_mean)
        return self._ds

    @property
    def graphs(self):
        from common import dataset_graphs

        if self._graphs is None:
            self._graphs = dataset_graphs(self.ds)
        return self._graphs

    def gaf_raw(self, r0, r1):
        return self.ds.gaf_raw(self.seed, r0, r1, self.params)


def wl_config1(args):
    import synth

    return Workload("configs[1]", f"single-species synthetic graph, {args.nodes} nodes, {args.haps} strain paths, short-read GAF",
                    SEED, [args.nodes], [args.haps], synth.GafParams())


def wl_config2():
    import synth

    return Workload("configs[2]", "HiFi long-read GAF (mean 15 kb) on a 100-species graph, 5 M nodes", 20261017 + 3, [50_000] * 100, [5] * 100,
                    synth.GafParams(long_reads=True, id_pair_suffix=False, p_secondary=0.1), backbone_mean=300.0)


def wl_config3():
    import synth

    return Workload("configs[3]", "multi-species synthetic graph: 1,000 species, 20 M nodes, 5,000 strain paths, short-read GAF", 20261017 + 4,
                    [20_000] * 1000, [5] * 1000, synth.GafParams())


def wl_config4():
    import synth

    return Workload("configs[4]", "large-scale stress: 50 M-node graph (1,000 species x 50 k nodes), 5,000 strain paths, short-read GAF", 20261017 + 5,
                    [50_000] * 1000, [5] * 1000, synth.GafParams())


def workload_name(args, n_gpus):
    return (f"BASELINE configs[1]: single-species synthetic graph, {args.nodes} nodes, {args.haps} strain paths, "
            f"{args.records} short-read GAF records per GPU x {n_gpus} GPU")


def make_oracle(wl):
    from oracle import cpu as ocpu

    o = ocpu.CpuOracle(0)
    o.set_ranges(wl.ds.ranges())
    for s, g in enumerate(wl.graphs):
        o.set_graph(s, g[0], g[1])
    t_prep = o.prepare_graphs()
    return o, t_prep


def oracle_results(o, n_species):
    out = {"counts": o.species_counts()}
    for s in range(n_species):
        if o.species_error(s):
            continue
        out[f"bases{s}"] = o.node_bases(s)
        out[f"cov{s}"] = o.node_cov(s)
        out[f"trio{s}"] = o.trio_bases(s)
        sc, sl = o.path_sums(s)
        out[f"pcov{s}"], out[f"plen{s}"] = sc, sl
        U, nz = o.hap_trio_counts(s)
        out[f"U{s}"], out[f"nz{s}"] = U, nz
    return out


def gpu_results(ctx, n_species):
    out = {"counts": ctx.species_counts()}
    for s in range(n_species):
        out[f"bases{s}"] = ctx.node_bases(s)
        out[f"cov{s}"] = ctx.node_cov(s)
        out[f"trio{s}"] = ctx.trio_bases(s)
        sc, sl = ctx.path_sums(s)
        out[f"pcov{s}"], out[f"plen{s}"] = sc, sl
        U, nz = ctx.hap_trio_counts(s)
        out[f"U{s}"], out[f"nz{s}"] = U, nz
    return out


def same_results(a, b):
    """Bit-exact equality of two result dictionaries (integer arrays); returns (equal, first differing key)."""
    if set(a) != set(b):
        return False, "keys: " + ",".join(sorted(set(a) ^ set(b))[:4])
    for k in sorted(a):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape or not np.array_equal(x.astype(np.int64), y.astype(np.int64)):
            return False, k
    return True, None


def digest(res):
    h = hashlib.sha256()
    for k in sorted(res):
        h.update(k.encode())
        h.update(np.ascontiguousarray(np.asarray(res[k]).astype(np.int64)).tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------------------
# reference arm: the CPU port of the reference's path, all host threads
# --------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    import synth

    wl = wl_config1(args)
    o, t_prep = make_oracle(wl)
    sample = min(args.records, args.cpu_sample)
    buf, nbytes = wl.gaf_raw(0, sample)
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
    whole = nrec >= args.records
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus),
                   "sample": (f"each step = all {nrec} records of rank 0's batch" if whole else f"each step = first {nrec} records of rank 0's batch"),
                   "graph_setup_s": t_prep},
        "cpu_baseline": {"value": value, "unit": "records/s", "cores": o.threads, "kind": "port",
                         "sample": f"{'all' if whole else 'first'} {nrec} records of rank 0's batch per step; C++ port of the reference "
                                   f"(oracle/oracle_cpu.cpp), {o.threads} threads; the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device (the library has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x: float, op: str) -> float:
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_obj(self, obj):
        if not self.dist:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def bcast_obj(self, obj):
        if not self.dist:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


def new_ctx(D, wl, reserve, with_comm):
    from pantax_b200 import api

    ctx = api.PantaxGpu(D.local_rank)
    ctx.set_ranges(wl.ds.ranges())
    t0 = time.perf_counter()
    for s, g in enumerate(wl.graphs):
        ctx.upload_graph(s, g[0], g[1])
    ctx.commit_graphs()
    t_graph = time.perf_counter() - t0
    ctx.graph_commit_s = ctx.stats().get("graph_commit_ms", 0.0) / 1e3
    if reserve:
        ctx.reserve(reserve)  # before comm_init: sizes the peer-memory id boxes
    if with_comm and D.world > 1:
        uid = D.bcast_obj(api.PantaxGpu.comm_unique_id() if D.rank == 0 else None)
        ctx.comm_init(D.world, D.rank, uid)
    return ctx, t_graph


PIECE_RECORDS = 8_000_000  # ~0.9 GB of short-read text per device buffer


def load_resident(ctx, wl, r0, r1, cudart):
    """Generates records [r0, r1) in pieces of <= ~1 GB and copies each straight into a device GAF buffer.
    Returns ([(buffer id, bytes)], total bytes, '>'/'<' count = walk nodes)."""
    import synth

    bufs, total, walk = [], 0, 0
    step = PIECE_RECORDS if not wl.params.long_reads else 1_000_000
    for a in range(r0, r1, step):
        b = min(r1, a + step)
        buf, nbytes = wl.gaf_raw(a, b)
        arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(nbytes,))
        walk += int(np.count_nonzero(arr == ord(">")) + np.count_nonzero(arr == ord("<")))
        bid, dptr = ctx.gaf_buffer_alloc(nbytes)
        assert cudart.cudaMemcpy(C.c_void_p(dptr), buf, C.c_size_t(nbytes), 1) == 0
        synth.lib().synth_free(buf)
        bufs.append((bid, nbytes))
        total += nbytes
    return bufs, total, walk


def step_resident(ctx, bufs):
    ctx.rewind()
    for bid, nbytes in bufs:
        ctx.ingest_gaf_device(bid, nbytes)
    ctx.finalize()


def stream_host(ctx, wl, r0, r1):
    """Host path: records [r0, r1) generated piece by piece and pushed through ptx_ingest_gaf (chunks split lines nowhere)."""
    import synth

    step = PIECE_RECORDS if not wl.params.long_reads else 1_000_000
    for a in range(r0, r1, step):
        b = min(r1, a + step)
        buf, nbytes = wl.gaf_raw(a, b)
        ctx.ingest_gaf(buf.value, nbytes, is_last=(b == r1))
        synth.lib().synth_free(buf)


def timed_steps(D, ctx, bufs, steps, warmup, sample_clocks):
    for _ in range(max(warmup, 3)):
        step_resident(ctx, bufs)
    n_rec = ctx.num_records
    _, _, launches0 = ctx.timing()
    sampler = ClockSampler(D.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    D.barrier()
    t0 = time.perf_counter()
    kern = []
    for _ in range(steps):
        step_resident(ctx, bufs)
        st = ctx.stats()  # CUDA-event times of this step's kernels (events on the library's stream)
        kern.append((st["count_ms"], st["ingest_ms"], st["finalize_ms"], st["ingest_launches"], st.get("apply_ms", 0.0)))
    D.barrier()
    dt = D.reduce(time.perf_counter() - t0, "max")
    clocks = sampler.stop() if sampler else None
    _, _, launches1 = ctx.timing()
    k = np.array(kern, dtype=np.float64)
    return dict(n_rec=n_rec, dt=dt, clocks=clocks, launches=int(launches1 - launches0), count_ms=float(k[:, 0].mean()),
                ingest_ms=float(k[:, 1].mean()), final_ms=float(k[:, 2].mean()), n_launch=max(1, int(k[-1, 3])), apply_ms=float(k[:, 4].mean()))


def roofline_block(t, nbytes, walk_nodes, p_hit, kernel_name, traffic):
    peak, peak_src = peaks()
    n_rec, n_launch = t["n_rec"], t["n_launch"]
    L = nbytes / n_rec
    W = walk_nodes / n_rec
    # B_rec = L + 12 W + 8 max(W-2,0) p_hit.  The path runs as two kernels: k_ingest reads the text
    # (L bytes per record), k_apply does the node/trio accumulation (the 12 W + 8 (W-2) p_hit part).
    b_apply = 12.0 * W + 8.0 * max(W - 2.0, 0.0) * p_hit
    b_rec = L + b_apply
    achieved = n_rec * L / max(t["ingest_ms"], 1e-9) / 1e6
    achieved_apply = n_rec * b_apply / max(t["apply_ms"], 1e-9) / 1e6
    achieved_path = n_rec * b_rec / max(t["ingest_ms"] + t["apply_ms"], 1e-9) / 1e6
    return {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_record": L, "mean_line_bytes": L, "mean_walk_nodes": W, "trio_hit_rate": p_hit,
            "kernel_ms": t["ingest_ms"] / n_launch, "launches_per_step": n_launch, "count_ms": t["count_ms"], "finalize_ms": t["final_ms"],
            "k_apply": {"kernel_ms": t["apply_ms"] / n_launch, "algorithmic_bytes_per_record": b_apply, "achieved": achieved_apply,
                        "frac": achieved_apply / peak},
            "whole_path": {"kernels": "k_ingest + k_apply", "algorithmic_bytes_per_record": b_rec, "achieved": achieved_path,
                           "frac": achieved_path / peak}}


def static_traffic(n_rec):
    for name in ("r2_ingest_traffic.json", "r1_ingest_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("records") == n_rec:
                    return tj.get("dram_bytes_per_launch")
            except Exception:
                pass
    return None


def cross_rank_check(D, res):
    """Every rank's reduced result must be rank 0's."""
    ds = D.gather_obj(digest(res))
    return all(d == ds[0] for d in ds)


def secondary_config(D, args, cudart, which="auto"):
    """One more BASELINE config, same step definition: N=1 -> configs[2] (HiFi), N>1 -> configs[3] (1,000 species, sharded);
    which="config4": configs[4] (50 M nodes, 5,000 paths, 1 B records over the ranks; run at N=8 or with --config4)."""
    import synth
    from pantax_b200 import api

    full_stream_check = True
    prefix_div = 64
    if which == "config4":
        wl, total = wl_config4(), args.config4_records or 1_000_000_000
        kernel_name = "k_ingest_s"
        full_stream_check = False  # streaming 1 B records through one GPU would take minutes: prefix parity + rank agreement only
        prefix_div = 512
    elif D.world == 1:
        wl, total = wl_config2(), args.secondary_records or 1_000_000
        kernel_name = "k_ingest_l (node-parallel walk decode)"
    else:
        wl, total = wl_config3(), args.secondary_records or 200_000_000
        kernel_name = "k_ingest_s"
    per = total // D.world
    r0, r1 = D.rank * per, (D.rank + 1) * per
    S = wl.ds.n_species
    ctx, t_graph = new_ctx(D, wl, per, with_comm=True)
    t0 = time.perf_counter()
    bufs, nbytes, walk = load_resident(ctx, wl, r0, r1, cudart)
    t_gen = time.perf_counter() - t0
    steps = 10 if (D.world == 1 and which == "auto") else 5
    t = timed_steps(D, ctx, bufs, steps, 3, sample_clocks=False)
    total_records = D.reduce(float(t["n_rec"]), "sum")
    res = gpu_results(ctx, S)
    ranks_equal = cross_rank_check(D, res)
    # parity: a prefix of the stream (all of it at N=1, 1/64 of the sharded 200 M) through a fresh single-GPU context and the C++ oracle
    prefix = total if (D.world == 1 and which == "auto") else max(total // prefix_div, 1)
    par = {"checked": False}
    p_hit = 0.0
    cpu = None
    if D.rank == 0:
        o, _ = make_oracle(wl)
        buf, nb = wl.gaf_raw(0, prefix)
        o.run(buf.value, nb)  # warm
        tc = time.perf_counter()
        o.run(buf.value, nb)
        cdt = time.perf_counter() - tc
        ws = o.workload_stats()
        p_hit = ws["trio_hits"] / max(1, ws["trio_windows"])
        c2, _ = new_ctx(D, wl, prefix, with_comm=False)
        c2.ingest_gaf(buf.value, nb, is_last=True)
        c2.finalize()
        eq, where = same_results(gpu_results(c2, S), oracle_results(o, S))
        synth.lib().synth_free(buf)
        par = {"checked": True, "equal": bool(eq), "first_difference": where, "what": f"first {o.n_records} records (a fresh single-GPU context) "
               f"vs the C++ oracle: species counts, bases, covered bases, trio bases, path sums, hap trio counts of all {S} species"}
        cpu = {"value": o.n_records / cdt, "unit": "records/s", "cores": o.threads, "kind": "port", "sample": f"first {o.n_records} records"}
        if D.world > 1 and full_stream_check:  # shard invariance at full size: one GPU streaming every rank's batch must give the reduced result
            c2.reset()
            stream_host(c2, wl, 0, per * D.world)
            c2.finalize()
            eq2, where2 = same_results(gpu_results(c2, S), res)
            par["sharded_equals_1gpu_streamed"] = bool(eq2)
            par["sharded_first_difference"] = where2
        c2.close()
        del o
    par["ranks_equal"] = bool(ranks_equal)
    rf = roofline_block(t, nbytes, walk, p_hit, kernel_name, None)
    out = {"workload": f"BASELINE {wl.key}: {wl.name}, {total} records over {D.world} GPU", "value": total_records * steps / t["dt"], "unit": "records/s",
           "steps": steps, "ms_per_step": 1e3 * t["dt"] / steps, "records_per_gpu": int(t["n_rec"]), "gaf_bytes_per_gpu": int(nbytes),
           "graph_setup_s": t_graph, "graph_commit_s": ctx.graph_commit_s, "generate_s": t_gen, "roofline": rf, "parity": par, "cpu_baseline": cpu, "gpu_launches": t["launches"]}
    ctx.close()
    return out


def run_ours(args):
    D = Dist()
    import synth
    from pantax_b200 import api

    world, rank = D.world, D.rank
    wl = wl_config1(args)
    R = args.records
    cudart = C.CDLL("libcudart.so")

    # one process per GPU: stay on the cores (and the memory) of the GPU's own NUMA node before anything is allocated
    numa_cpus = api.bind_host_thread_to_gpu(D.local_rank) if world > 1 else None

    # ---- synthetic inputs: graph replicated, rank r owns records [r*R, (r+1)*R)
    buf, nbytes = wl.gaf_raw(rank * R, (rank + 1) * R)
    pinned = api.PinnedBuffer(nbytes)
    C.memmove(pinned.ptr, buf, nbytes)
    synth.lib().synth_free(buf)
    text = pinned.array
    walk_nodes = int(np.count_nonzero(text == ord(">")) + np.count_nonzero(text == ord("<")))

    ctx, t_graph = new_ctx(D, wl, R, with_comm=True)
    bid, dptr = ctx.gaf_buffer_alloc(nbytes)
    assert cudart.cudaMemcpy(C.c_void_p(dptr), C.c_void_p(pinned.ptr), C.c_size_t(nbytes), 1) == 0
    bufs = [(bid, nbytes)]

    # ---- value: text resident in HBM
    t = timed_steps(D, ctx, bufs, args.steps, args.warmup, sample_clocks=True)
    n_rec = t["n_rec"]
    total_records = D.reduce(float(n_rec), "sum")
    value = total_records * args.steps / t["dt"]
    res_resident = gpu_results(ctx, 1)  # what the LAST timed step left in the context
    parity = {"checked": True, "ranks_equal": bool(cross_rank_check(D, res_resident))}

    # ---- cpu baseline + parity of the timed configuration itself, rank 0 (N=1: against the oracle on the same records)
    cpu = None
    p_hit = 0.0
    if rank == 0 and not args.no_cpu_baseline:
        o, _ = make_oracle(wl)
        sample = min(R, args.cpu_sample) if world == 1 else min(R, 2_000_000)
        sbuf, sbytes = wl.gaf_raw(0, sample)
        o.run(sbuf.value, sbytes)  # warm
        t0 = time.perf_counter()
        o.run(sbuf.value, sbytes)
        cdt = time.perf_counter() - t0
        ws = o.workload_stats()
        p_hit = ws["trio_hits"] / max(1, ws["trio_windows"])
        whole = o.n_records == n_rec
        cpu = {"value": o.n_records / cdt, "unit": "records/s", "cores": o.threads, "kind": "port",
               "sample": f"{'all' if whole else 'first'} {o.n_records} of {n_rec} records; oracle/oracle_cpu.cpp (C++ port of the reference, "
                         f"{o.threads} threads); stages s: parse+classify+counts {o.times()[0]:.3f}, id grouping {o.times()[1]:.3f}, "
                         f"coverage {o.times()[2]:.3f}, stats {o.times()[3]:.3f}"}
        if world == 1:
            if whole:  # the timed context itself against the oracle, all records
                eq, where = same_results(res_resident, oracle_results(o, 1))
                what = f"timed context after the last step vs the C++ oracle on all {o.n_records} records"
            else:
                c2, _ = new_ctx(D, wl, sample, with_comm=False)
                c2.ingest_gaf(sbuf.value, sbytes, is_last=True)
                c2.finalize()
                eq, where = same_results(gpu_results(c2, 1), oracle_results(o, 1))
                c2.close()
                what = f"first {o.n_records} records through a fresh context vs the C++ oracle"
            parity.update({"equal": bool(eq), "first_difference": where,
                           "what": what + ": species counts, bases, covered bases, trio bases, path sums, hap trio counts"})
        synth.lib().synth_free(sbuf)
        del o
    if world > 1:
        # shard invariance: rank 0 streams EVERY rank's batch through one single-GPU context (host path, 64 MB pieces);
        # integer sums and ORs are order-free, so the NCCL-reduced result must be identical
        eq = True
        where = None
        if rank == 0:
            c2, _ = new_ctx(D, wl, R * world, with_comm=False)
            stream_host(c2, wl, 0, R * world)
            c2.finalize()
            eq, where = same_results(gpu_results(c2, 1), res_resident)
            c2.close()
        eq = D.bcast_obj(bool(eq))
        parity.update({"equal": bool(eq) and parity["ranks_equal"], "first_difference": where,
                       "what": f"NCCL-reduced result of {world} ranks vs one GPU streaming all {R * world} records through the host path; every rank's result vs rank 0's"})
    p_hit = D.bcast_obj(p_hit)
    roofline = roofline_block(t, nbytes, walk_nodes, p_hit, "k_ingest_s (GAF parse -> record table + CSR walks)", static_traffic(n_rec))

    # ---- e2e: host buffers through the C ABI, H2D inside the timed region, results read back
    e2e = None
    if not args.no_e2e:
        ctx.reset()  # releases the resident buffer
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
        e_steps = min(args.steps, 20)
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            out = step_e2e()
        D.barrier()
        edt = D.reduce(time.perf_counter() - t0, "max")
        e2e = {"value": total_records * e_steps / edt, "unit": "records/s", "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * edt / e_steps, "steps": e_steps}
        # chunk invariance at full size: the host path cut the text into 64 MB pieces, the resident path was one chunk
        keys = ["counts", "bases0", "cov0", "trio0", "pcov0", "plen0", "U0", "nz0"]
        eq_e2e, where_e2e = same_results({k: np.array(v) for k, v in zip(keys, out)}, res_resident)
        parity["resident_equals_chunked_host_path"] = bool(eq_e2e)
        if not eq_e2e:
            parity["chunked_first_difference"] = where_e2e

    graph_commit_main = ctx.graph_commit_s
    secondary = None
    tertiary = None
    if not args.no_secondary:
        t_graph_main = t_graph
        n_trios_main = ctx.n_trios(0)
        ctx.close()
        pinned.free()
        secondary = secondary_config(D, args, cudart)
        if args.config4 or world == 8:
            tertiary = secondary_config(D, args, cudart, which="config4")
    else:
        t_graph_main = t_graph
        n_trios_main = ctx.n_trios(0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * t["dt"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "records_per_gpu": int(n_rec), "gaf_bytes_per_gpu": int(nbytes),
                       "l2": f"input text {nbytes / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "graph_setup_s": t_graph_main, "graph_commit_s": graph_commit_main, "unique_trios": n_trios_main,
                       "host_cores_rank0": (f"{len(numa_cpus)} cores of the GPU's NUMA node" if numa_cpus else "not restricted"), "timing": "wall clock over K steps between "
                       "barrier+synchronize, max over ranks; kernel times from CUDA events on the library stream"},
            "clocks": t["clocks"], "e2e": e2e, "gpu_launches": t["launches"], "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "secondary": secondary,
        }
        if tertiary is not None:
            line["tertiary"] = tertiary
        print(json.dumps(line), flush=True)
    if D.dist:
        D.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
