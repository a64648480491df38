"""CPU oracle for PanTax's alignment-to-abundance hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a deliberately naive, pure-Python restatement of the reference's
algorithm (LuoGroup2023/PanTax v2.1.0, `pantax/src/*.rs`).  It exists to CHECK
the CUDA path; nothing in `pantax_b200/` may import it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg use `oracle/`.

PARITY UNPINNED: the reference ships no golden vectors, fixtures or asserting
tests for this path, its Rust toolchain and crates are absent
from this image so it cannot be run here, and the path's third-party pieces
(polars 0.46.0 CSV reader, regex 1.12.2, hashbrown/fxhash iteration order,
nalgebra 0.33.2 f32 gemv) are un-vendored.  The pins we do have are the
hand-derived known-answer tests of SURVEY.md section 8c (tests/golden/kat_*.json), the
two real GAF lines quoted in the reference's comments (profile.rs:396,
profile.rs:813) and a second, independent restatement in C++
(oracle/oracle_cpu.cpp) cross-checked against this one on randomised inputs.

Every function cites the reference lines it follows (paths relative to
/root/reference/pantax/src/).

Conventions shared by both oracles and the CUDA library (documented in
DESIGN.md "GAF dialect"):
  * lines are split on b'\\n'; one trailing b'\\r' is dropped; empty lines and
    lines starting with b'@' are skipped (rcls.rs:123 comment prefix);
  * fields are split on b'\\t', no quoting (rcls.rs:125); a field equal to b'*'
    is null in every column except the read id (rcls.rs:124); a missing field
    is null;
  * integer columns (2,7,8,9,12) are `[+-]?[0-9]{1,18}` over the whole field,
    anything else is null (polars' strict=false cast, rcls.rs:132-134);
  * node ids are maximal ASCII digit runs of the path field (regex `\\d+`,
    rcls.rs:455; `-?\\d+` at profile.rs:769 - a '-' never occurs in a GAF walk
    of numeric segment ids, and is treated as a separator here); runs longer
    than 18 digits are dropped (rcls.rs:244 `filter_map(parse().ok())`).
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

_INT_RE = re.compile(rb"^[+-]?[0-9]{1,18}$")
_DIGITS_RE = re.compile(rb"[0-9]+")

UNCLASSIFIED = "U"


# --------------------------------------------------------------------------
# a1  GAF text -> rows                                   rcls.rs:119-146
# --------------------------------------------------------------------------
@dataclass
class Row:
    read_id: bytes
    read_len: Optional[int]
    path: Optional[bytes]
    read_path_len: Optional[int]
    read_start: Optional[int]
    read_end: Optional[int]
    mapq: Optional[int]
    species: str = UNCLASSIFIED
    line: bytes = b""


def _int_or_null(f: Optional[bytes]) -> Optional[int]:
    if f is None or not _INT_RE.match(f):
        return None
    return int(f)


def _field(fields: List[bytes], k: int) -> Optional[bytes]:
    """1-based column k; '*' and missing are null (rcls.rs:124)."""
    if k > len(fields):
        return None
    f = fields[k - 1]
    return None if f == b"*" else f


def iter_gaf_lines(data: bytes) -> Iterable[bytes]:
    for line in data.split(b"\n"):
        if line.endswith(b"\r"):
            line = line[:-1]
        if not line or line.startswith(b"@"):
            continue
        yield line


def load_gaf(data: bytes) -> List[Row]:
    """rcls.rs:119-146 `load_gaf_file_lazy`: columns 1,2,6,7,8,9,12."""
    rows = []
    for line in iter_gaf_lines(data):
        f = line.split(b"\t")
        rows.append(
            Row(
                read_id=f[0],
                read_len=_int_or_null(_field(f, 2)),
                path=_field(f, 6),
                read_path_len=_int_or_null(_field(f, 7)),
                read_start=_int_or_null(_field(f, 8)),
                read_end=_int_or_null(_field(f, 9)),
                mapq=_int_or_null(_field(f, 12)),
                line=line,
            )
        )
    return rows


# --------------------------------------------------------------------------
# a2  read classification                               rcls.rs:237-258, 306-323
# --------------------------------------------------------------------------
def digit_runs(path: Optional[bytes]) -> List[int]:
    if path is None:  # null path behaves as "" (rcls.rs:311 into_no_null_iter)
        return []
    return [int(m) for m in _DIGITS_RE.findall(path) if len(m) <= 18]


def classify(path: Optional[bytes], ranges: Sequence[Tuple[str, int, int]]) -> str:
    """rcls.rs:237-258: first range in FILE ORDER with min>=start && max<=end."""
    nodes = digit_runs(path)
    if not nodes:
        lo, hi = -1, -1
    else:
        lo, hi = min(nodes), max(nodes)
    for name, start, end in ranges:
        if lo >= start and hi <= end:
            return name
    return UNCLASSIFIED


def rcls_profile(data: bytes, ranges: Sequence[Tuple[str, int, int]], species_column: Optional[Sequence[str]] = None) -> List[Row]:
    """rcls.rs:452-458.  With `species_column` (strain-only resume, profile.rs:3367-3385): the species of row i is
    species_column[i], the third column of reads_classification.tsv, hstacked onto the GAF columns.  A row whose
    walk leaves the node range of its supplied species is treated as "U" (the reference would index outside the
    species graph); `rows.label_out_of_range` reports it."""
    rows = load_gaf(data)
    if species_column is None:
        for r in rows:
            r.species = classify(r.path, ranges)
        return rows
    if len(species_column) < len(rows):
        raise ValueError("species column shorter than the GAF")
    bounds = {name: (s, e) for name, s, e in ranges}
    bad = False
    for r, sp in zip(rows, species_column):
        if sp != UNCLASSIFIED and r.path is not None:
            ids = digit_runs(r.path)
            if ids and (min(ids) < bounds[sp][0] or max(ids) > bounds[sp][1]):
                sp, bad = UNCLASSIFIED, True
        r.species = sp
    rows = RowList(rows)
    rows.label_out_of_range = bad
    return rows


class RowList(list):
    label_out_of_range = False


# --------------------------------------------------------------------------
# a3  species level counts + abundance                  profile.rs:208-349
# --------------------------------------------------------------------------
def species_counts(rows: Sequence[Row]) -> "OrderedDict[str, List[int]]":
    """Integer part of profile.rs:208-297 over rows with species != U.

    Returns species -> [read_count, sum_read_len, less_multi, uniq_count]
    (less_multi = #(3<=mapq<=60), uniq_count = #(mapq==60)); insertion order =
    first appearance.
    """
    out: "OrderedDict[str, List[int]]" = OrderedDict()
    for r in rows:
        if r.species == UNCLASSIFIED:
            continue
        c = out.setdefault(r.species, [0, 0, 0, 0])
        c[0] += 1
        c[1] += r.read_len if r.read_len is not None else 0
        if r.mapq is not None and 3 <= r.mapq <= 60:
            c[2] += 1
            if r.mapq == 60:
                c[3] += 1
    return out


def equal_length_test(rows: Sequence[Row]) -> Tuple[bool, Optional[int]]:
    """profile.rs:311-322: distinct read_len among the first 1000 non-U rows == 1."""
    first = [r.read_len for r in rows if r.species != UNCLASSIFIED][:1000]
    uniq = []
    for v in first:
        if v not in uniq:
            uniq.append(v)
    if len(uniq) == 1:
        return True, uniq[0]
    return False, None


def species_profiling(
    rows: Sequence[Row], species_len: Dict[str, float], filtered: bool = True
) -> List[Tuple[str, float, float]]:
    """profile.rs:299-349.  Returns [(taxid, predicted_abundance, predicted_coverage)]
    sorted by abundance descending (ties: unspecified in the reference; here stable)."""
    counts = species_counts(rows)
    equal, read_len0 = equal_length_test(rows)
    base = OrderedDict()
    for sp, (rc, sl, lm, uq) in counts.items():
        if filtered and not (uq > 0 and lm > rc / 10.0):  # profile.rs:239-245
            continue
        base[sp] = rc * read_len0 if equal else sl  # :246 / :291
    absolute = OrderedDict()
    for sp, b in base.items():
        ln = species_len.get(sp)
        absolute[sp] = (b / ln) if ln is not None else float("nan")  # left join, :333-337
    total = sum(v for v in absolute.values())
    table = [(sp, v / total, v) for sp, v in absolute.items()]
    table.sort(key=lambda t: -t[1])
    return table


# --------------------------------------------------------------------------
# a4  read grouping / duplicate-id rule                 profile.rs:361-463
# --------------------------------------------------------------------------
@dataclass
class Record:
    read_id: bytes
    path: bytes
    read_path_len: int
    read_start: int
    read_end: int
    species: str


def group_reads_by_species(rows: Sequence[Row]) -> Dict[str, List[Record]]:
    """profile.rs:361-463 on rows with species != U (profile.rs:3352-3356).

    Uniqueness is tested over ALL non-U rows (:376), rows with a null
    path/c7/c8/c9 are dropped from the records (:380-399).  If any id repeats,
    records are grouped by id and a group survives only if all its records
    share one species (:406-437).
    """
    seen = set()
    unique = True
    records: List[Record] = []
    for r in rows:
        if r.species == UNCLASSIFIED:
            continue
        if r.read_id in seen:
            unique = False
        seen.add(r.read_id)
        if None in (r.path, r.read_path_len, r.read_start, r.read_end):
            continue
        records.append(Record(r.read_id, r.path, r.read_path_len, r.read_start, r.read_end, r.species))
    out: Dict[str, List[Record]] = {}
    if unique:
        for rec in records:
            out.setdefault(rec.species, []).append(rec)
        return out
    groups: "OrderedDict[bytes, List[Record]]" = OrderedDict()
    for rec in records:
        groups.setdefault(rec.read_id, []).append(rec)
    for _id, grp in groups.items():
        if len({g.species for g in grp}) == 1:
            out.setdefault(grp[0].species, []).extend(grp)
    return out


# --------------------------------------------------------------------------
# a5  graph                                              types.rs:51-55
# --------------------------------------------------------------------------
@dataclass
class Graph:
    nodes_len: List[int]
    paths: "OrderedDict[str, List[int]]" = field(default_factory=OrderedDict)  # sorted by name (BTreeMap)

    def sorted_paths(self) -> List[Tuple[str, List[int]]]:
        return sorted(self.paths.items(), key=lambda kv: kv[0].encode())


def read_gfa(text: str, previous: int = 0) -> Graph:
    """profile.rs:466-545."""
    nodes_len: List[int] = []
    paths: Dict[str, List[int]] = {}
    idx = 0
    for line in text.split("\n"):
        if line.startswith("S"):
            parts = line.strip().split("\t")
            if len(parts) < 3:
                continue
            nid = int(parts[1]) - 1 - previous
            if nid != idx:
                raise ValueError("Node ID out of order or mismatch")  # :489
            idx += 1
            if len(parts[2]) == 0:
                raise ValueError("Node length 0 appears in the GFA")  # :494
            nodes_len.append(len(parts[2]))
        elif line.startswith("W") or line.startswith("P"):
            parts = line.strip().split("\t")
            if parts[0] == "W":
                hap = parts[1]
                p = [int(m) - 1 - previous for m in re.findall(r"-?\d+", parts[-1])]
            else:
                hap = parts[1].split("#")[0]
                p = [int(m) - 1 - previous for m in re.findall(r"\d+", parts[2] if len(parts) > 2 else "")]
            paths.setdefault(hap, []).extend(p)  # :540
    g = Graph(nodes_len)
    g.paths = OrderedDict(sorted(paths.items(), key=lambda kv: kv[0].encode()))
    return g


def read_and_zip_gfa(text: str, previous: int = 0):
    """zip.rs:78-171: read_gfa plus what the database step does before it serialises the Graph - a W line whose walk starts with
    `<` and a P line whose first step ends with `-` are stored reversed (:124, :137, :147-149); returns (Graph, min + 1, max + 1,
    is_pan) = the row `zip` appends to the range file (:316-327)."""
    nodes_len: List[int] = []
    paths: Dict[str, List[int]] = {}
    mins: List[int] = []
    maxs: List[int] = []
    idx = 0
    for line in text.split("\n"):
        if line.startswith("S"):
            parts = line.strip().split("\t")
            if len(parts) < 3:
                continue
            if int(parts[1]) - 1 - previous != idx:
                raise ValueError("Node ID out of order or mismatch")
            idx += 1
            if len(parts[2]) == 0:
                raise ValueError("Node length 0 appears in the GFA")
            nodes_len.append(len(parts[2]))
        elif line.startswith("W") or line.startswith("P"):
            parts = line.strip().split("\t")
            if parts[0] == "W":
                hap = parts[1]
                rev = parts[-1].startswith("<")
                p = [int(m) - 1 - previous for m in re.findall(r"-?\d+", parts[-1])]
            else:
                hap = parts[1].split("#")[0]
                fld = parts[2] if len(parts) > 2 else ""
                rev = fld.split(",")[0].endswith("-")
                p = [int(m) - 1 - previous for m in re.findall(r"\d+", fld)]
            if rev:
                p.reverse()
            mins.append(min(p))  # :151 unwrap: an empty path is a panic
            maxs.append(max(p))
            paths.setdefault(hap, []).extend(p)
    g = Graph(nodes_len)
    g.paths = OrderedDict(sorted(paths.items(), key=lambda kv: kv[0].encode()))
    return g, min(mins) + 1, max(maxs) + 1, 1 if len(paths) > 1 else 0


# --------------------------------------------------------------------------
# a6  unique trio nodes                                  profile.rs:658-740
# --------------------------------------------------------------------------
def canon(a: int, b: int, c: int) -> Tuple[int, int, int]:
    return (c, b, a) if a > c else (a, b, c)  # :672-678


def trio_nodes_info(graph: Graph):
    """profile.rs:658-740.

    Returns (trio_map: canonical key -> idx, trio_len[idx], owner_hap[idx]).
    The reference numbers unique trios in FxHashSet iteration order (:684,
    :711), which is unpinned; both oracles and the CUDA library use the
    deterministic order (owner hap index in name order, window position).
    """
    count: Dict[Tuple[int, int, int], int] = {}
    for _name, path in graph.sorted_paths():
        for i in range(len(path) - 2):
            k = canon(path[i], path[i + 1], path[i + 2])
            count[k] = count.get(k, 0) + 1  # :689-702 multiplicity across all paths
    trio_map: Dict[Tuple[int, int, int], int] = {}
    trio_len: List[int] = []
    owner: List[int] = []
    for h, (_name, path) in enumerate(graph.sorted_paths()):
        for i in range(len(path) - 2):
            k = canon(path[i], path[i + 1], path[i + 2])
            if count[k] == 1:  # :709
                trio_map[k] = len(trio_len)
                trio_len.append(graph.nodes_len[k[0]] + graph.nodes_len[k[1]] + graph.nodes_len[k[2]])  # :712
                owner.append(h)
    return trio_map, trio_len, owner


def trio_nodes_info_reference_order(graph: Graph):
    """profile.rs:658-740 with the reference's OWN numbering: unique trios in the iteration order of the
    FxHashSet of :659-685 (oracle/fx_hashset.py restates fxhash 0.2.1 + std's hashbrown table).  Same return shape as
    trio_nodes_info; only the index order differs - visible in the f64 sums of first_filter_paths (:1123-1146)."""
    from . import fx_hashset

    paths = [p for _name, p in graph.sorted_paths()]
    _all, uniq = fx_hashset.reference_trio_numbering(paths)
    owner_of: Dict[Tuple[int, int, int], int] = {}
    for h, path in enumerate(paths):
        for i in range(len(path) - 2):
            owner_of[canon(path[i], path[i + 1], path[i + 2])] = h  # unique trios have exactly one owner
    trio_map = {k: i for i, k in enumerate(uniq)}
    trio_len = [graph.nodes_len[k[0]] + graph.nodes_len[k[1]] + graph.nodes_len[k[2]] for k in uniq]
    owner = [owner_of[k] for k in uniq]
    return trio_map, trio_len, owner


# --------------------------------------------------------------------------
# a7  node coverage                                      profile.rs:743-1026
# --------------------------------------------------------------------------
class StartBeyondNode(Exception):
    """profile.rs:854 `assert!(read.read_start <= node_len)` (a panic -> abort)."""


def get_node_abundances(
    nodes_len: Sequence[int],
    trio_map: Dict[Tuple[int, int, int], int],
    trio_len: Sequence[int],
    range_start: int,
    records: Sequence[Record],
):
    """profile.rs:743-1026.  `range_start` is the species' 1-based first global id
    (local id = global - range_start, profile.rs:2886 + :790).

    Returns (bases[N] int, trio_bases[T] int, node_base_cov[N] int,
             node_abundance[N] float, trio_abundance[T] float).
    """
    n = len(nodes_len)
    bases = [0] * n
    trio_bases = [0] * len(trio_len)
    bits = [bytearray(l) for l in nodes_len]  # :776-781 one byte per base

    for rd in records:
        nodes = [m - range_start for m in digit_runs(rd.path)]  # :788-792
        if not nodes:
            continue  # :794
        target = rd.read_end - rd.read_start  # :800
        seen = 0
        rl: Dict[int, int] = {nd: 0 for nd in nodes}  # :806-808
        undup = set()
        ps, pe = rd.read_start, rd.read_end
        if len(nodes) == 1:  # :811
            nd = nodes[0]
            if target < 0:
                continue  # :821-827 (skips the trio part too; <3 nodes anyway)
            rl[nd] += target
            bases[nd] += target  # :829
            if ps < pe and pe <= nodes_len[nd]:  # :832
                if ps >= 0:  # `ps as usize` of a negative wraps -> empty range
                    for j in range(ps, pe):
                        bits[nd][j] = 1
        else:
            for i, nd in enumerate(nodes):
                ln = nodes_len[nd]
                if i == 0:
                    if ps > ln:
                        raise StartBeyondNode(f"read start is bigger than node len: {ps} > {ln}")  # :854
                    aln, sidx = ln - ps, ps  # :856
                elif i == len(nodes) - 1:
                    if target < seen:
                        target = seen  # :858
                    aln, sidx = target - seen, 0
                else:
                    aln, sidx = ln, 0  # :861
                if sidx >= 0:  # negative start wraps in `as usize` -> empty range
                    for j in range(sidx, min(sidx + aln, ln)):  # :871
                        bits[nd][j] = 1
                seen += aln  # :878
                if nd not in undup:  # :879-882 first occurrence only
                    undup.add(nd)
                    rl[nd] += aln
                    bases[nd] += aln
        if len(nodes) < 3:
            continue  # :886
        for i in range(len(nodes) - 2):
            a, b, c = nodes[i], nodes[i + 1], nodes[i + 2]
            s = rl[a] + rl[b] + rl[c]  # :897-900
            t = trio_map.get((a, b, c))
            if t is None:
                t = trio_map.get((c, b, a))  # :902-904
            if t is not None:
                trio_bases[t] += s
    cov = [sum(b) for b in bits]  # :844/:874 -> :1018-1023
    node_ab = [bases[i] / nodes_len[i] for i in range(n)]  # :980-990
    trio_ab = [trio_bases[i] / trio_len[i] for i in range(len(trio_len))]  # :1006-1016
    return bases, trio_bases, cov, node_ab, trio_ab


# --------------------------------------------------------------------------
# a8  strain statistics                                  profile.rs:1028-1227
# --------------------------------------------------------------------------
def seq_sum(xs) -> float:
    """`iter().sum::<f64>()`: one addition per element in order (builtin sum() is compensated since Python 3.12)."""
    s = 0.0
    for x in xs:
        s += x
    return s


def zscore_filter(data: Sequence[float], threshold: float = 3.0) -> List[float]:
    """profile.rs:1028-1051 (population sigma; sigma==0 -> empty)."""
    if not data:
        return []
    mean = seq_sum(data) / len(data)
    std = math.sqrt(seq_sum((x - mean) * (x - mean) for x in data) / len(data))
    if std == 0.0:
        return []
    return [x for x in data if abs((x - mean) / std) < threshold]


def hap_trio_counts(n_haps: int, owner: Sequence[int], trio_bases: Sequence[int]):
    """Integer part of profile.rs:1112-1135: U_h = #unique trios owned by hap h,
    nz_h = # of those with abundance > 0 (== trio_bases > 0 since lengths > 0)."""
    U = [0] * n_haps
    nz = [0] * n_haps
    for t, h in enumerate(owner):
        U[h] += 1
        if trio_bases[t] > 0:
            nz[h] += 1
    return U, nz


def first_filter_paths(
    graph: Graph,
    owner: Sequence[int],
    trio_ab: Sequence[float],
    node_ab_opt: Sequence[float],
    fr: float,
    shift: bool = False,
):
    """profile.rs:1080-1227.  Returns (possible_paths_idx, metrics) with
    metrics[h] = dict(unique_trio_nodes_fraction, frequencies_mean)."""
    paths = graph.sorted_paths()
    H = len(paths)
    T = len(owner)
    metrics = [dict(unique_trio_nodes_fraction=None, frequencies_mean=None) for _ in range(H)]
    possible: List[int] = []
    same_path = False
    if H != 1 and T * H != 0:  # hap2trio_nodes_m.len() = T*H (:1094)
        for h in range(H):
            idxs = [t for t in range(T) if owner[t] == h]  # :1114-1116
            if not idxs:
                continue  # :1119
            nzf = [trio_ab[t] for t in idxs if trio_ab[t] > 0.0]  # :1123-1133 (trio idx order)
            frac = len(nzf) / len(idxs)  # :1135
            metrics[h]["unique_trio_nodes_fraction"] = round_half_away(frac * 100.0) / 100.0
            zf = zscore_filter(nzf, 3.0)
            fmean = (seq_sum(zf) / len(zf)) if zf else 0.0
            if shift:  # :1140-1165
                if fmean >= 1.0:
                    thr = min(fr + (0.8 - fr) * fmean / 100.0, 0.8)
                else:
                    thr = fr * fmean
                if frac < thr:
                    continue
            else:  # :1166-1181
                if frac < fr:
                    continue
            metrics[h]["frequencies_mean"] = fmean
            possible.append(h)
    elif H != 1:  # T == 0  (:1187-1209)
        vals = [p for _n, p in paths]
        if all(v == vals[0] for v in vals[1:]):
            same_path = True
            nzf = [x for x in node_ab_opt if x > 0.0]
            fmean = (seq_sum(nzf) / len(nzf)) if nzf else 0.0
            metrics[0]["frequencies_mean"] = round_half_away(fmean * 100.0) / 100.0
            possible.append(0)
        else:
            possible = list(range(H))
    else:  # H == 1  (:1211-1225)
        nzf = [x for x in node_ab_opt if x > 0.0]
        fmean = (seq_sum(nzf) / len(nzf)) if nzf else 0.0
        metrics[0]["frequencies_mean"] = round_half_away(fmean * 100.0) / 100.0
        possible.append(0)
    return possible, metrics, same_path


def round_half_away(x: float) -> float:
    """Rust f64::round (half away from zero)."""
    if math.isnan(x) or math.isinf(x):
        return x
    t = float(math.trunc(x))
    return t + math.copysign(1.0, x) if abs(x - t) >= 0.5 else t


# --------------------------------------------------------------------------
# a9  path covered fraction                              profile.rs:2705-2729
# --------------------------------------------------------------------------
def path_sums(graph: Graph, node_base_cov: Sequence[int]):
    """Integer sums over the DISTINCT nodes of each path (binary incidence,
    profile.rs:2706-2711): returns (sum_cov[h], sum_len[h]) for every path in
    name order.  The reference then forms the ratio in f32 (f64 in cbc_opt,
    profile.rs:1952-1977)."""
    sc, sl = [], []
    for _name, p in graph.sorted_paths():
        d = set(p)
        sc.append(sum(node_base_cov[v] for v in d))
        sl.append(sum(graph.nodes_len[v] for v in d))
    return sc, sl


def path_cov_ratio(graph: Graph, node_base_cov: Sequence[int], f32: bool = True) -> List[float]:
    """profile.rs:2705-2729 (highs_opt; the same block in gurobi/cplex/glpk, f64 in cbc_opt :1952-1977):
    `RowDVector<f32>(cov) * DMatrix<f32>(incidence)` and the same with the node lengths, then component_div.
    nalgebra 0.33 multiplies a 1 x nvert row by an nvert x npaths matrix with one gemv per output column (a
    dimension of 1 keeps it off the matrixmultiply kernels), and gemv accumulates `y = 1*a[v]*x[v] + 1*y` over
    v = 0..nvert in index order: a SEQUENTIAL f32 sum over the distinct nodes of the path in node-index order
    (nodes outside the path add 0).  Once a running sum passes 2^24 it is no longer the exact integer."""
    import numpy as np

    ft = np.float32 if f32 else np.float64
    out = []
    for _name, p in graph.sorted_paths():
        acc_c, acc_l = ft(0), ft(0)
        for v in sorted(set(p)):
            acc_c = ft(acc_c + ft(node_base_cov[v]))
            acc_l = ft(acc_l + ft(graph.nodes_len[v]))
        out.append(float(ft(acc_c / acc_l)) if len(p) else float("nan"))
    return out


# --------------------------------------------------------------------------
# a11  long-read best-alignment filter                   gaf_filter.rs:22-97
# --------------------------------------------------------------------------
def _parse_i32(b: bytes) -> Optional[int]:
    if not re.match(rb"^[+-]?[0-9]+$", b):
        return None
    v = int(b)
    return v if -(2 ** 31) <= v < 2 ** 31 else None


def _parse_f64(b: bytes) -> Optional[float]:
    try:
        s = b.decode("ascii")
        if not re.match(r"^[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?$", s):
            return None
        return float(s)
    except Exception:
        return None


def filter_max_alignment(data: bytes) -> List[bytes]:
    """gaf_filter.rs:44-97.  Deterministic tie rule (the reference's is a race,
    :80-93): the FIRST qualifying line in file order is kept per read id, and
    output order is file order."""
    recs = []
    for line in data.split(b"\n"):
        if line.endswith(b"\r"):
            line = line[:-1]
        f = line.strip().split(b"\t")  # :23 `line.trim().split('\t')`
        if len(f) < 16:
            continue
        m = _parse_i32(f[9])
        ident = _parse_f64(f[15].rsplit(b":", 1)[-1])
        q = _parse_i32(f[11])
        a, b_ = _parse_i32(f[3]), _parse_i32(f[2])
        if None in (m, ident, q, a, b_):
            continue
        span = ((a - b_ + 2 ** 31) % 2 ** 32) - 2 ** 31  # i32 subtraction: release builds wrap (Cargo.toml has no overflow-checks)
        recs.append((line, f[0], m, ident, q, span))
    best: Dict[bytes, Tuple[int, float]] = {}
    for _l, rid, m, ident, _q, _s in recs:
        e = best.get(rid)
        if e is None or m > e[0] or (m == e[0] and ident > e[1]):
            best[rid] = (m, ident)
    out, written = [], set()
    for line, rid, m, ident, q, span in recs:
        if q > 20 and span > 1000 and best[rid] == (m, ident) and rid not in written:
            written.add(rid)
            out.append(line)
    return out


# --------------------------------------------------------------------------
# whole path, per species                               profile.rs:2884-2944, 3291-3323
# --------------------------------------------------------------------------
def coverage_all_species(
    data: bytes,
    ranges: Sequence[Tuple[str, int, int]],
    graphs: Dict[str, Graph],
    species_column: Optional[Sequence[str]] = None,
):
    """rcls -> grouping -> per-species trio table + node coverage + path sums.

    Returns (rows, counts, per_species) with per_species[taxid] = dict(bases,
    trio_bases, cov, trio_len, owner, U, nz, sum_cov, sum_len, error)."""
    rows = rcls_profile(data, ranges, species_column)
    counts = species_counts(rows)
    grouped = group_reads_by_species(rows)
    start_of = {name: s for name, s, _e in ranges}
    out = {}
    for sp, g in graphs.items():
        trio_map, trio_len, owner = trio_nodes_info(g)
        recs = grouped.get(sp, [])
        try:
            bases, trio_bases, cov, node_ab, trio_ab = get_node_abundances(
                g.nodes_len, trio_map, trio_len, start_of[sp], recs
            )
            err = None
        except StartBeyondNode as e:
            out[sp] = dict(error=str(e))
            continue
        U, nz = hap_trio_counts(len(g.paths), owner, trio_bases)
        sc, sl = path_sums(g, cov)
        out[sp] = dict(
            bases=bases, trio_bases=trio_bases, cov=cov, trio_len=trio_len, owner=owner,
            trio_keys=sorted(trio_map, key=trio_map.get), U=U, nz=nz, sum_cov=sc, sum_len=sl,
            node_ab=node_ab, trio_ab=trio_ab, error=err,
        )
    return rows, counts, out


# --------------------------------------------------------------------------
# a10 / f2  PAO model rows, second filter, abundance constraint
#           profile.rs:2699-2813, 1229-1285, 3028-3070
# --------------------------------------------------------------------------
def highs_rows_dense(graph: Graph, possible_paths_idx: Sequence[int], node_abundance: Sequence[float], minimization_min_cov: float = 0.0,
                     fixed_zero: Sequence[int] = ()):
    """The RowProblem highs_opt builds (profile.rs:2699-2813), literally: a dense nvert x npaths f32 incidence matrix
    (`coeff_matrix[(v, pos_idx)] = 1.0`), columns x_i (0..1.05 max), binary indicators, one y per node with depth > 0
    (objective 1/n_y); rows in the order of the `pb.add_row` calls.  Returns (c, A, lo, hi, lb, ub, integer) as numpy arrays.
    (No node sampling: callers stay below --sample.)"""
    import numpy as np

    paths = [p for _n, p in graph.sorted_paths()]
    nvert, npaths = len(node_abundance), len(possible_paths_idx)
    max_val = max(node_abundance) if len(node_abundance) else float("-inf")
    coeff = np.zeros((nvert, npaths), dtype=np.float32)
    for i, path in enumerate(paths):
        if i in possible_paths_idx:
            pos = list(possible_paths_idx).index(i)
            for v in path:
                coeff[v, pos] = 1.0
    valid = [v for v, ab in enumerate(node_abundance) if ab > 0.0]
    n_y = len(valid)
    ncol = 2 * npaths + n_y
    c = np.zeros(ncol)
    lb = np.zeros(ncol)
    ub = np.full(ncol, np.inf)
    integer = np.zeros(ncol, dtype=np.int64)
    for i in range(npaths):
        ub[i] = 1.05 * max_val
        ub[npaths + i] = 1.0
        integer[npaths + i] = 1
    for k in range(n_y):
        c[2 * npaths + k] = 1.0 / n_y
    rows, lo, hi = [], [], []

    def add_row(terms, rlo, rhi):
        r = np.zeros(ncol)
        for col, val in terms:
            r[col] = val
        rows.append(r)
        lo.append(rlo)
        hi.append(rhi)

    for i in range(npaths):
        add_row([(npaths + i, 1.0), (i, -1.0 / (2.0 * max_val))], -(minimization_min_cov / (2.0 * max_val)), np.inf)
    add_row([(npaths + i, 1.0) for i in range(npaths)], -np.inf, float(npaths))
    for k, v in enumerate(valid):
        coeffs = [(j, float(val)) for j, val in enumerate(coeff[v]) if abs(val) > 1e-12]
        add_row(coeffs + [(2 * npaths + k, -1.0)], -np.inf, node_abundance[v])
        add_row(coeffs + [(2 * npaths + k, 1.0)], node_abundance[v], np.inf)
    for i in fixed_zero:
        add_row([(i, 1.0)], 0.0, 0.0)
    return c, np.array(rows), np.array(lo), np.array(hi), lb, ub, integer


def second_filter_paths(metrics: List[dict], possible_paths_idx: Sequence[int], orign_n_haps: int, hap2trio_nodes_m_size: int, same_path_flag: bool,
                        fc: float = 0.46, sr: float = 0.85):
    """profile.rs:1229-1285 on a list of dicts with the HapMetrics fields; returns (second_opt, second_possible_paths_idx)."""
    second_opt, keep = False, []
    if orign_n_haps != 1 and hap2trio_nodes_m_size > 0:
        second_opt = True
        for idx in possible_paths_idx:
            m = metrics[idx]
            fmean = m.get("frequencies_mean") or 0.0
            if fmean == 0.0:
                continue
            sol = m["first_sol"]
            f = abs(sol - fmean) / (sol + fmean)
            fr = round_half_away(f * 100.0) / 100.0
            m["divergence"] = fr
            if fr > fc:
                if fr <= 0.6:
                    if m["unique_trio_nodes_fraction"] * m["path_cov_ratio"] < sr or sol == 0.0:
                        continue
                    m["is_rescue"] = True
                    keep.append(idx)
                else:
                    continue
            elif fr <= fc and sol != 0.0:
                keep.append(idx)
    elif (orign_n_haps != 1 and hap2trio_nodes_m_size == 0 and same_path_flag) or orign_n_haps == 1:
        m = metrics[0]
        if m["frequencies_mean"] > 0.0:
            sol = m["first_sol"]
            f = abs(sol - m["frequencies_mean"]) / (sol + m["frequencies_mean"])
            m["divergence"] = round_half_away(f * 100.0) / 100.0
            m["second_sol"] = sol
    elif orign_n_haps != 1 and hap2trio_nodes_m_size == 0 and not same_path_flag:
        for idx in possible_paths_idx:
            metrics[idx]["second_sol"] = metrics[idx].get("first_sol")
    return second_opt, keep


def abundace_constraint(species_coverage: float, metrics: List[dict]):
    """profile.rs:3028-3070."""
    absab = []
    for m in metrics:
        if m.get("is_rescue") is True and m.get("first_sol") is not None and m.get("second_sol") is not None:
            m["second_sol"] = min(m["first_sol"], m["second_sol"])
        absab.append(m["second_sol"] if m.get("second_sol") is not None else 0.0)
    total = 0.0
    for v in absab:
        total += v
    diff = abs(total - species_coverage) / ((total + species_coverage) / 2.0)
    for m in metrics:
        m["total_cov_diff"] = diff
    if absab and max(absab) > 1.05 * species_coverage:
        factor = species_coverage / total
        for m in metrics:
            if not (m.get("is_rescue") or False) and m.get("second_sol") is not None:
                m["second_sol"] = m["second_sol"] * factor
