"""Host tail of strain profiling over the GPU outputs: what `optimize_otu` does after `get_node_abundances`
(profile.rs:2936-3026) and `abundance_est` (profile.rs:3091-3289), so that a GAF goes all the way to
`strain_abundance.txt` without the Rust binary.  Names follow the reference (paths relative to
/root/reference/pantax/src):

  zscore_filter          profile.rs:1028-1051
  first_filter_paths     profile.rs:1080-1227   from ptx_hap_trio_counts / ptx_trio_depth / ptx_trio_table(owner)
  build_pao_model        profile.rs:2735-2813   the PAO ILP as a CSR constraint matrix (the reference fills a dense
                                                nvert x npaths f32 matrix, which does not exist at 20 M nodes)
  highs_opt              profile.rs:2689-2882   two solves with HiGHS (scipy.optimize.milp wraps the same solver)
  second_filter_paths    profile.rs:1229-1285
  abundace_constraint    profile.rs:3028-3070   (the reference's spelling)
  abundance_est          profile.rs:3091-3289   strain_abundance.txt (11 columns) + ori_strain_abundance.txt

Stage 4 of north_star - the ILP itself - stays in a host solver; this module only hands it the GPU's numbers in the
layout the reference builds, and carries the post-solver rules.  Everything numeric that depends on the reads comes from
libpantax_gpu.so (`pantax_b200.api`); the CPU restatements used by the tests are never imported here.

`sample_sorted` (profile.rs:1287-1295: rand 0.9 `StdRng(42)` + `choose_multiple` when a species has more than `--sample`
= 500 000 covered nodes) draws through `rand09.py`, a restatement of that generator and sampler; float columns are printed in
ryu's format (what polars' CSV writer uses) by `fmt_f64`.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import api


@dataclass
class ProfilingArgs:
    """The options of ProfilingConfig (types.rs:57-91) this stage reads, with the CLI defaults (cli.rs / main.rs:102-171)."""
    unique_trio_nodes_fraction: float = 0.3      # --fr (0.5 for long reads)
    unique_trio_nodes_mean_count_f: float = 0.46  # --fc
    single_cov_ratio: float = 0.85               # --sr
    single_cov_diff: float = 0.2                 # --sd
    min_cov: float = 0.0
    min_depth: float = 0.0
    minimization_min_cov: float = 0.0
    sample_nodes: int = 500_000                  # --sample
    sample_test: bool = False
    shift: bool = False
    full: bool = True                            # main.rs:160: hard-wired
    gurobi_threads: int = 1


@dataclass
class HapMetrics:  # profile.rs:1065-1078
    otu: Optional[str] = None
    hap_id: Optional[str] = None
    unique_trio_nodes_fraction: Optional[float] = None
    frequencies_mean: Optional[float] = None
    path_cov_ratio: Optional[float] = None
    first_sol: Optional[float] = None
    divergence: Optional[float] = None
    second_sol: Optional[float] = None
    is_rescue: Optional[bool] = None
    total_cov_diff: Optional[float] = None


@dataclass
class OptVar:  # GurobiOptVar, profile.rs:1053-1063
    otu: str
    hap_metrics: List[HapMetrics]
    possible_paths_idx: List[int] = field(default_factory=list)
    second_possible_paths_idx: List[int] = field(default_factory=list)
    orign_n_haps: int = 0
    hap2trio_nodes_m_size: int = 0
    same_path_flag: bool = False
    second_opt: bool = False


def rust_round(x: float) -> float:
    """f64::round: half away from zero.  (Not floor(x + 0.5): that addition rounds, e.g. 0.49999999999999994 + 0.5 == 1.0.)"""
    if math.isnan(x) or math.isinf(x):
        return x
    t = float(math.trunc(x))
    return t + math.copysign(1.0, x) if abs(x - t) >= 0.5 else t  # x - trunc(x) is exact


def seq_sum(xs) -> float:
    """Iterator::sum::<f64>() of the reference: one f64 addition per element, in order.  (Python's builtin sum() is
    Neumaier-compensated since 3.12 and np.sum is pairwise - neither gives the reference's bits.)"""
    s = 0.0
    for x in xs:
        s += x
    return s


def zscore_filter(data: Sequence[float], threshold: float = 3.0) -> List[float]:
    """profile.rs:1028-1051: population mean / standard deviation; std == 0 -> nothing survives."""
    n = len(data)
    if n == 0:
        return []
    mean = seq_sum(data) / n
    var = seq_sum((x - mean) * (x - mean) for x in data) / n
    std = math.sqrt(var)
    if std == 0.0:
        return []
    return [x for x in data if abs((x - mean) / std) < threshold]


def first_filter_paths(opt: OptVar, hap_names: Sequence[str], paths: Sequence[np.ndarray], trio_owner: np.ndarray, trio_depth: np.ndarray,
                       node_depth_opt: np.ndarray, args: ProfilingArgs, trio_order: Optional[np.ndarray] = None) -> None:
    """profile.rs:1080-1227.  The dense trio x hap matrix of the reference is the `owner` column of ptx_trio_table here
    (a unique trio belongs to exactly one hap).  `trio_order` (api.trio_ref_order) puts the trios into the reference's own
    numbering first, so that the f64 sums below add a hap's abundances in the reference's order; without it they are visited
    in the library's order (hap, position) - same sets, same counts, sums equal up to the rounding of a reordered f64 sum."""
    if trio_order is not None:
        trio_owner = np.asarray(trio_owner)[trio_order]
        trio_depth = np.asarray(trio_depth)[trio_order]
    for i, h in enumerate(hap_names):
        opt.hap_metrics[i].otu = opt.otu
        opt.hap_metrics[i].hap_id = h
    H, T = len(hap_names), len(trio_owner)
    opt.orign_n_haps = H
    opt.hap2trio_nodes_m_size = H * T
    if H != 1 and opt.hap2trio_nodes_m_size != 0:
        for hap_idx in range(H):
            mine = trio_depth[trio_owner == hap_idx]  # index order
            if len(mine) == 0:
                continue
            nz = [float(x) for x in mine if x > 0.0]
            fraction = len(nz) / len(mine)
            opt.hap_metrics[hap_idx].unique_trio_nodes_fraction = rust_round(fraction * 100.0) / 100.0
            kept = zscore_filter(nz, 3.0)
            fmean = seq_sum(kept) / len(kept) if kept else 0.0
            if args.shift:
                if fmean >= 1.0:
                    thr = min(args.unique_trio_nodes_fraction + (0.8 - args.unique_trio_nodes_fraction) * fmean / 100.0, 0.8)
                else:
                    thr = args.unique_trio_nodes_fraction * fmean
            else:
                thr = args.unique_trio_nodes_fraction
            if fraction < thr:
                continue
            opt.hap_metrics[hap_idx].frequencies_mean = fmean
            opt.possible_paths_idx.append(hap_idx)
    elif H != 1:
        first = np.asarray(paths[0])
        all_same = all(len(p) == len(first) and np.array_equal(np.asarray(p), first) for p in paths[1:])
        if all_same:
            opt.same_path_flag = True
            nzv = node_depth_opt[node_depth_opt > 0.0]
            fmean = float(np.add.accumulate(nzv)[-1] / len(nzv)) if len(nzv) else 0.0
            opt.hap_metrics[0].frequencies_mean = rust_round(fmean * 100.0) / 100.0
            opt.possible_paths_idx.append(0)
        else:
            opt.possible_paths_idx = list(range(H))
    else:
        nzv = node_depth_opt[node_depth_opt > 0.0]
        fmean = float(np.add.accumulate(nzv)[-1] / len(nzv)) if len(nzv) else 0.0
        opt.hap_metrics[0].frequencies_mean = rust_round(fmean * 100.0) / 100.0
        opt.possible_paths_idx.append(0)


def sample_sorted(valid_nodes: np.ndarray, sample_size: int, seed: int) -> np.ndarray:
    """profile.rs:1287-1295: `StdRng::seed_from_u64(seed)` + `choose_multiple` + `sort_unstable`, with rand 0.9.2's generator and
    index sampler restated in rand09.py - the same seed draws the same nodes as the reference (unpinned: restated, not run)."""
    from . import rand09

    return rand09.choose_multiple_sorted(np.asarray(valid_nodes), sample_size, seed)


@dataclass
class PaoModel:
    """min c.x  s.t.  lo <= A x <= hi, bounds, integrality - the RowProblem of profile.rs:2754-2813 in CSR."""
    c: np.ndarray
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray
    lo: np.ndarray
    hi: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    integrality: np.ndarray
    npaths: int
    nodes: np.ndarray  # the sampled covered nodes, one y variable each

    def dense(self) -> np.ndarray:
        a = np.zeros((len(self.lo), len(self.c)))
        for r in range(len(self.lo)):
            a[r, self.indices[self.indptr[r]:self.indptr[r + 1]]] = self.data[self.indptr[r]:self.indptr[r + 1]]
        return a


def build_pao_model(paths: Sequence[np.ndarray], possible_paths_idx: Sequence[int], node_abundance: np.ndarray, args: ProfilingArgs,
                    fixed_zero: Sequence[int] = ()) -> PaoModel:
    """profile.rs:2699-2813 (highs_opt): variables x_0..x_{P-1} (path abundances, 0 <= x <= 1.05 max depth), P binary indicators,
    one y per covered node (objective 1/n each); rows: indicator_i - x_i/(2 max) >= -min_cov/(2 max); sum indicators <= P;
    for every covered node v: sum_{p contains v} x_p - y_v <= depth_v and sum x_p + y_v >= depth_v.  `fixed_zero`: positions
    (into possible_paths_idx) whose x is pinned to 0 for the second solve (profile.rs:2846-2850).
    The incidence comes straight from the path node lists (CSR by node), never as a dense nvert x npaths matrix."""
    P = len(possible_paths_idx)
    max_val = float(np.max(node_abundance)) if len(node_abundance) else float("-inf")
    valid = np.nonzero(node_abundance > 0.0)[0]
    limit = 500 if args.sample_test else (args.sample_nodes if args.sample_nodes > 0 else 0)
    if limit and len(valid) > limit:
        valid = sample_sorted(valid, limit, 42)
    n_y = len(valid)
    # node -> positions of the candidate paths that contain it (binary incidence: a node visited twice counts once)
    pos_of_node = np.full(len(node_abundance), -1, dtype=np.int64)
    pos_of_node[valid] = np.arange(n_y)
    per_row: List[List[int]] = [[] for _ in range(n_y)]
    for j, p_idx in enumerate(possible_paths_idx):
        for r in np.unique(pos_of_node[np.unique(np.asarray(paths[p_idx], dtype=np.int64))]):
            if r >= 0:
                per_row[int(r)].append(j)
    ncol = 2 * P + n_y
    c = np.zeros(ncol)
    c[2 * P:] = 1.0 / n_y if n_y else 0.0
    lb = np.zeros(ncol)
    ub = np.full(ncol, np.inf)
    ub[:P] = 1.05 * max_val
    ub[P:2 * P] = 1.0
    integrality = np.zeros(ncol, dtype=np.int64)
    integrality[P:2 * P] = 1
    indptr, indices, data, lo, hi = [0], [], [], [], []

    def add_row(cols, vals, rlo, rhi):
        indices.extend(cols)
        data.extend(vals)
        indptr.append(len(indices))
        lo.append(rlo)
        hi.append(rhi)

    for i in range(P):
        add_row([P + i, i], [1.0, -1.0 / (2.0 * max_val)], -(args.minimization_min_cov / (2.0 * max_val)), np.inf)
    add_row(list(range(P, 2 * P)), [1.0] * P, -np.inf, float(P))
    for r, v in enumerate(valid):
        cols = per_row[r]
        add_row(cols + [2 * P + r], [1.0] * len(cols) + [-1.0], -np.inf, float(node_abundance[v]))
        add_row(cols + [2 * P + r], [1.0] * len(cols) + [1.0], float(node_abundance[v]), np.inf)
    for i in fixed_zero:
        add_row([i], [1.0], 0.0, 0.0)
    return PaoModel(c, np.array(indptr, dtype=np.int64), np.array(indices, dtype=np.int64), np.array(data), np.array(lo), np.array(hi),
                    lb, ub, integrality, P, valid)


def solve_pao(model: PaoModel) -> np.ndarray:
    """HiGHS through scipy.optimize.milp (the `highs` crate of the reference drives the same solver); status must be optimal
    (profile.rs:2824, :2862).  Returns the solution vector (x, indicators, y)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix

    a = csr_matrix((model.data, model.indices, model.indptr), shape=(len(model.lo), len(model.c)))
    res = milp(model.c, constraints=LinearConstraint(a, model.lo, model.hi), bounds=Bounds(model.lb, model.ub), integrality=model.integrality)
    if res.status != 0 or res.x is None:
        raise RuntimeError(f"HiGHS did not reach optimality: {res.message}")
    return res.x


def second_filter_paths(opt: OptVar, args: ProfilingArgs) -> None:
    """profile.rs:1229-1285."""
    if opt.orign_n_haps != 1 and opt.hap2trio_nodes_m_size > 0:
        opt.second_opt = True
        keep = []
        for idx in opt.possible_paths_idx:
            m = opt.hap_metrics[idx]
            fmean = m.frequencies_mean if m.frequencies_mean is not None else 0.0
            if fmean == 0.0:
                continue
            sol = m.first_sol
            f = abs(sol - fmean) / (sol + fmean)
            f_rounded = rust_round(f * 100.0) / 100.0
            m.divergence = f_rounded
            if f_rounded > args.unique_trio_nodes_mean_count_f:
                if f_rounded <= 0.6:
                    ratio = m.unique_trio_nodes_fraction * m.path_cov_ratio
                    if ratio < args.single_cov_ratio or sol == 0.0:
                        continue
                    m.is_rescue = True
                    keep.append(idx)
            elif sol != 0.0:
                keep.append(idx)
        opt.second_possible_paths_idx = keep
    elif (opt.orign_n_haps != 1 and opt.hap2trio_nodes_m_size == 0 and opt.same_path_flag) or opt.orign_n_haps == 1:
        m = opt.hap_metrics[0]
        if m.frequencies_mean is not None and m.frequencies_mean > 0.0:
            sol = m.first_sol
            f = abs(sol - m.frequencies_mean) / (sol + m.frequencies_mean)
            m.divergence = rust_round(f * 100.0) / 100.0
            m.second_sol = sol
    else:
        for idx in opt.possible_paths_idx:
            opt.hap_metrics[idx].second_sol = opt.hap_metrics[idx].first_sol


def highs_opt(opt: OptVar, paths: Sequence[np.ndarray], node_abundance: np.ndarray, path_cov_ratio: np.ndarray, args: ProfilingArgs) -> None:
    """profile.rs:2689-2882: path_cov_ratio, first solve, second filter, second solve with the filtered paths pinned to 0."""
    for k, idx in enumerate(opt.possible_paths_idx):
        opt.hap_metrics[idx].path_cov_ratio = float(path_cov_ratio[idx])
    model = build_pao_model(paths, opt.possible_paths_idx, node_abundance, args)
    x = solve_pao(model)
    for k, idx in enumerate(opt.possible_paths_idx):
        opt.hap_metrics[idx].first_sol = float(x[k])
    second_filter_paths(opt, args)
    if not opt.second_opt:
        return
    pinned = [k for k, idx in enumerate(opt.possible_paths_idx) if idx not in opt.second_possible_paths_idx]
    model2 = build_pao_model(paths, opt.possible_paths_idx, node_abundance, args, fixed_zero=pinned)
    x2 = solve_pao(model2)
    # profile.rs:2866-2880: sols2 = the first len(second_possible_paths_idx) columns, zipped with possible_paths_idx in order
    sols2 = x2[:min(len(x2), len(opt.second_possible_paths_idx))]
    for idx, sol in zip(opt.possible_paths_idx, sols2):
        if idx in opt.second_possible_paths_idx:
            opt.hap_metrics[idx].second_sol = float(sol)


def optimize_otu(ctx: "api.PantaxGpu", species: int, otu: str, nodes_len: np.ndarray, paths: Sequence[np.ndarray], hap_names: Sequence[str],
                 args: ProfilingArgs) -> List[HapMetrics]:
    """profile.rs:2884-3026 after load_from_zip_graph: every read-dependent number comes from the GPU context."""
    node_depth, trio_depth, _cov = api.get_node_abundances(ctx, species)
    keys, _tlen, owner = api.trio_nodes_info(ctx, species)
    node_depth_opt = np.where(node_depth > args.min_depth, node_depth, 0.0)
    opt = OptVar(otu=otu, hap_metrics=[HapMetrics() for _ in hap_names])
    order = api.trio_ref_order(paths, keys) if len(hap_names) > 1 and len(owner) else None
    first_filter_paths(opt, hap_names, paths, owner, trio_depth, node_depth_opt, args, trio_order=order)
    if opt.possible_paths_idx:
        ratio = api.path_cov_ratio(ctx, species, paths, nodes_len, f32=True)
        highs_opt(opt, paths, node_depth, ratio, args)
    return opt.hap_metrics


def abundace_constraint(species_coverage: float, metrics: List[HapMetrics]) -> None:
    """profile.rs:3028-3070; `species_coverage` = predicted_coverage of the species in species_abundance.txt."""
    absab = []
    for m in metrics:
        if m.is_rescue is True and m.first_sol is not None and m.second_sol is not None:
            m.second_sol = min(m.first_sol, m.second_sol)
        absab.append(m.second_sol if m.second_sol is not None else 0.0)
    total = 0.0
    for v in absab:
        total += v
    diff = abs(total - species_coverage) / ((total + species_coverage) / 2.0) if (total + species_coverage) != 0 else float("nan")
    for m in metrics:
        m.total_cov_diff = diff
    if absab and max(absab) > 1.05 * species_coverage:
        factor = species_coverage / total
        for m in metrics:
            if not (m.is_rescue or False) and m.second_sol is not None:
                m.second_sol = m.second_sol * factor


def fmt_f64(v: Optional[float]) -> str:
    """A float as ryu prints it (polars' CSV writer): shortest round-trip digits, `1.0` for integers, plain decimals for
    exponents in (-5, 16), otherwise `1.5e-7` / `1e16` (no `+`, no zero padding).  None -> empty field."""
    if v is None:
        return ""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    r = repr(float(v))
    sign = "-" if r.startswith("-") else ""
    r = r.lstrip("-")
    if "e" in r:
        mant, ex = r.split("e")
        ex = int(ex)
    else:
        mant, ex = r, 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    point = len(ip) + ex - (len(ip + fp) - len((ip + fp).lstrip("0")))  # position of the decimal point relative to `digits`
    digits = digits.rstrip("0") or "0"
    k = point  # value = 0.digits * 10^k
    if 0 < k <= 16:
        if len(digits) <= k:
            return sign + digits + "0" * (k - len(digits)) + ".0"
        return sign + digits[:k] + "." + digits[k:]
    if -5 < k <= 0:
        return sign + "0." + "0" * (-k) + digits
    e = k - 1
    return sign + digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + "e" + str(e)


def hap_id_of_genome(path_id: str) -> str:
    """profile.rs:3106-3145: file stem of the `id` column, first two `_`-separated pieces when it has an underscore."""
    stem = os.path.basename(path_id)
    stem = stem[: stem.rfind(".")] if "." in stem else stem
    return "_".join(stem.split("_")[:2]) if stem.count("_") >= 1 else stem


def abundance_est(args: ProfilingArgs, metrics: Sequence[HapMetrics], genomes_info: Sequence[Tuple[str, str, str, str, str]], out_path: str,
                  ori_path: Optional[str] = None) -> List[List[str]]:
    """profile.rs:3091-3289: joins the hap metrics with genomes_info.txt rows (genome_ID, strain_taxid, species_taxid, organism_name, id),
    predicted_abundance = coverage / sum, the two filters, sort by abundance (descending, stable), 11-column TSV.  Returns the rows."""
    by_hap: Dict[str, List[Tuple[str, str]]] = {}
    for gid, strain, _sp, _name, pid in genomes_info:
        by_hap.setdefault(hap_id_of_genome(pid), []).append((gid, strain))
    header = ["species_taxid", "strain_taxid", "genome_ID", "predicted_coverage", "predicted_abundance", "path_base_cov", "unique_trio_fraction",
              "uniq_trio_cov_mean", "first_sol", "strain_cov_diff", "total_cov_diff"]
    # :3176-3183 left join on hap_id: one row per matching genomes_info line (none -> one row with null genome columns)
    merged: List[Tuple[HapMetrics, Optional[str], Optional[str]]] = []
    for m in metrics:
        for gid, strain in by_hap.get(m.hap_id, [(None, None)]):
            merged.append((m, gid, strain))
    cov_sum = 0.0
    for m, _g, _s in merged:
        if m.second_sol is not None:
            cov_sum += m.second_sol

    def row(e: Tuple[HapMetrics, Optional[str], Optional[str]], total: float):
        m, gid, strain = e
        ab = None if m.second_sol is None else (m.second_sol / total if total != 0 else float("nan"))
        return [m.otu, strain or "", gid or "", fmt_f64(m.second_sol), fmt_f64(ab), fmt_f64(m.path_cov_ratio), fmt_f64(m.unique_trio_nodes_fraction),
                fmt_f64(m.frequencies_mean), fmt_f64(m.first_sol), fmt_f64(m.divergence), fmt_f64(m.total_cov_diff)]

    if ori_path:
        with open(ori_path, "w") as f:
            f.write("\t".join(header) + "\n")
            for e in merged:
                f.write("\t".join(row(e, cov_sum)) + "\n")
    group: Dict[str, int] = {}
    for m, _g, _s in merged:
        if m.hap_id is not None:  # count() skips nulls
            group[m.otu] = group.get(m.otu, 0) + 1
    kept = [e for e in merged
            if (group.get(e[0].otu, 0) > 1 or (e[0].total_cov_diff is not None and e[0].total_cov_diff <= args.single_cov_diff))
            and e[0].second_sol is not None and e[0].second_sol >= args.min_cov and e[0].second_sol != 0.0]
    total = 0.0
    for m, _g, _s in kept:
        total += m.second_sol
    kept.sort(key=lambda e: -(e[0].second_sol / total))
    rows = [row(e, total) for e in kept]
    with open(out_path, "w") as f:
        f.write("\t".join(header) + "\n")
        for r in rows:
            f.write("\t".join(r) + "\n")
    return rows


# ---------------------------------------------------------------------------------------------------------------------
# File-level strain stage: the tables pantax-gpu-profile wrote (GPU numbers) -> strain_abundance.txt
# ---------------------------------------------------------------------------------------------------------------------
def read_bin_graph(path: str) -> Tuple[np.ndarray, List[str], List[np.ndarray]]:
    """zip.rs:236-247 for a plain `.bin`: bincode 1.3 (little endian, fixed ints) of types.rs:51-55 - u64 n, i64 nodes_len[n], u64 n_paths,
    per path in key order: u64 klen, key, u64 plen, u64 node[plen].  Returns (nodes_len, hap names, paths as local node ids)."""
    raw = np.fromfile(path, dtype=np.uint8)
    pos = 0

    def u64() -> int:
        nonlocal pos
        if pos + 8 > len(raw):
            raise ValueError(f"{path}: truncated bincode graph")
        v = int(raw[pos:pos + 8].view("<u8")[0])
        pos += 8
        return v

    def take(nbytes: int) -> np.ndarray:
        nonlocal pos
        if nbytes < 0 or pos + nbytes > len(raw):
            raise ValueError(f"{path}: truncated bincode graph")
        v = raw[pos:pos + nbytes]
        pos += nbytes
        return v

    n = u64()
    nodes_len = take(8 * n).view("<i8").astype(np.int64)
    names, paths = [], []
    for _ in range(u64()):
        names.append(bytes(take(u64())).decode())
        paths.append(take(8 * u64()).view("<u8").astype(np.int64))
    if pos != len(raw):
        raise ValueError(f"{path}: {len(raw) - pos} bytes behind the graph")
    return nodes_len, names, paths


def _tsv(path: str) -> List[List[str]]:
    with open(path) as f:
        return [l.split("\t") for l in f.read().split("\n")[1:] if l]


def load_strain_inputs(wd: str, db: str, taxid: str, args: ProfilingArgs):
    """One species as pantax-gpu-profile left it in <wd>/strain_inputs: the Graph, node depths, and the state of GurobiOptVar after
    first_filter_paths (profile.rs:2936-2967).  Returns (opt, paths, node_depth, path_cov_ratio) or None when no node was covered."""
    si = os.path.join(wd, "strain_inputs")
    own = os.path.join(wd, "strain_graphs", f"{taxid}.bin")  # written by the driver when the db holds .bin.lz4 / .bin.zst / a GFA
    nodes_len, names, paths = read_bin_graph(own if os.path.exists(own) else os.path.join(db, "species_graph_info", f"{taxid}.bin"))
    node_rows = _tsv(os.path.join(si, f"{taxid}.nodes.tsv"))
    if not node_rows:
        return None  # no read of this species survived: the reference has no read cluster for it and skips optimize_otu (profile.rs:3300-3302)
    node_depth = np.zeros(len(nodes_len), dtype=np.float64)
    for r in node_rows:
        node_depth[int(r[0])] = float(r[2])
    rows = _tsv(os.path.join(si, f"{taxid}.paths.tsv"))
    if [r[0] for r in rows] != names:
        raise ValueError(f"{taxid}: paths.tsv and the graph list different haplotypes")
    H = len(rows)
    T = sum(int(r[1]) for r in rows)
    opt = OptVar(otu=taxid, hap_metrics=[HapMetrics(otu=taxid, hap_id=r[0]) for r in rows])
    opt.orign_n_haps = H
    opt.hap2trio_nodes_m_size = H * T
    ratio = np.full(H, np.nan)
    for h, r in enumerate(rows):
        m = opt.hap_metrics[h]
        if r[3] != "":
            m.unique_trio_nodes_fraction = float(r[3])
        if r[5] != "":
            ratio[h] = float(r[5])
        if r[8] == "1":
            opt.possible_paths_idx.append(h)
            if r[4] != "":
                m.frequencies_mean = float(r[4])
    if H > 1 and T == 0:
        opt.same_path_flag = all(len(p) == len(paths[0]) and np.array_equal(p, paths[0]) for p in paths[1:])
    return opt, paths, node_depth, ratio


def run_strain_stage(db: str, wd: str, args: ProfilingArgs, ori_path: Optional[str] = None) -> List[List[str]]:
    """profile.rs:3291-3323 (strain_profiling) from the files of a pantax-gpu-profile run: per species (species_range.txt order)
    highs_opt + abundace_constraint, then abundance_est -> <wd>/strain_abundance.txt.  Everything read-dependent in those files was
    computed on the GPU; this stage is the host solver and the joins."""
    species_cov: Dict[str, float] = {}
    for r in _tsv(os.path.join(wd, "species_abundance.txt")):
        species_cov[r[0]] = float(r[2])
    with open(os.path.join(db, "species_range.txt")) as f:
        order = [l.split("\t")[0] for l in f.read().split("\n") if l]
    metrics: List[HapMetrics] = []
    for taxid in order:
        if not os.path.exists(os.path.join(wd, "strain_inputs", f"{taxid}.paths.tsv")):
            continue
        loaded = load_strain_inputs(wd, db, taxid, args)
        if loaded is None:
            continue
        opt, paths, node_depth, ratio = loaded
        if opt.possible_paths_idx:
            highs_opt(opt, paths, node_depth, ratio, args)
        abundace_constraint(species_cov[taxid], opt.hap_metrics)
        metrics.extend(opt.hap_metrics)
    info = [tuple((r + [""] * 5)[:5]) for r in _tsv(os.path.join(db, "genomes_info.txt"))]
    return abundance_est(args, metrics, info, os.path.join(wd, "strain_abundance.txt"), ori_path)


def main(argv: Optional[Sequence[str]] = None) -> int:
    import argparse

    ap = argparse.ArgumentParser(prog="python -m pantax_b200.strain_tail",
                                 description="Strain stage after `pantax-gpu-profile --species --strain`: strain_inputs/ -> strain_abundance.txt")
    ap.add_argument("--db", "-d", required=True)
    ap.add_argument("--wd", "-T", default=".")
    ap.add_argument("--fr", type=float, default=None, help="fstrain (default 0.3, long reads 0.5)")
    ap.add_argument("--fc", type=float, default=0.46, help="dstrain")
    ap.add_argument("--sr", type=float, default=0.85)
    ap.add_argument("--sd", type=float, default=0.2)
    ap.add_argument("--long-read", action="store_true")
    ap.add_argument("--shift", action="store_true", help="must equal the --shift given to pantax-gpu-profile (first filter)")
    ap.add_argument("--min_cov", type=float, default=0.0)
    ap.add_argument("--min_depth", type=float, default=0.0)
    ap.add_argument("--sample", type=int, default=500_000)
    ap.add_argument("--sample-test", action="store_true")
    ap.add_argument("--ori", default="ori_strain_abundance.txt", help="profile.rs:3217 writes it to the current directory")
    a = ap.parse_args(argv)
    args = ProfilingArgs(unique_trio_nodes_fraction=a.fr if a.fr is not None else (0.5 if a.long_read else 0.3), unique_trio_nodes_mean_count_f=a.fc,
                         single_cov_ratio=a.sr, single_cov_diff=a.sd, min_cov=a.min_cov, min_depth=a.min_depth, sample_nodes=a.sample,
                         sample_test=a.sample_test, shift=a.shift)
    rows = run_strain_stage(a.db, a.wd, args, a.ori)
    print(f"- Strain level profiling: {len(rows)} strains written to {os.path.join(a.wd, 'strain_abundance.txt')}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
