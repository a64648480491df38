"""Byte-level fuzzing of the GAF dialect on the CPU: lines assembled from hostile fragments (signs, '*', empty
fields, CR, huge or zero-padded numbers, separators other than <>, missing columns, repeated tabs, comments) must be
read identically by the naive Python oracle, the C++ oracle and the code the kernels run (ptx_core.cuh through
tests/hostcheck.cpp, with a normal window, a tiny window that forces the re-parse path, and without the stash)."""
import numpy as np
import pytest

from common import assert_cpu_matches_py
from test_core_host import assert_hostcheck_matches

RANGES = [("a", 1, 60), ("b", 61, 100), ("c", 101, 104)]
GRAPHS = [
    (np.array([7, 3, 9, 1, 12, 5, 8, 2, 6, 4] * 6, dtype=np.int64),
     [np.arange(60, dtype=np.uint64), np.array([5, 4, 3, 4, 5, 6, 7, 20, 21, 22], dtype=np.uint64), np.arange(59, 30, -1).astype(np.uint64)],
     ["h1", "h2", "h3"]),
    (np.array([4, 10, 2, 30] * 10, dtype=np.int64), [np.arange(40, dtype=np.uint64), np.arange(0, 40, 3).astype(np.uint64)], ["k1", "k2"]),
    None,   # a species without a graph: classified and counted, never covered
]

IDS = [b"r%d", b"read/%d", b"*", b"S0R%d/1", b"S0R%d/2", b"x y %d", b"dup", b"dup2"]
INTS = [b"0", b"1", b"3", b"7", b"12", b"60", b"150", b"-1", b"+5", b"007", b"*", b"", b"1x", b"x", b" 4", b"4 ", b"999999999999999999",
        b"1000000000000000000", b"9223372036854775807", b"-", b"+", b"1e3", b"0x10", b"3.0"]
SEPS = [b">", b"<", b"-", b"_", b":", b"chr", b" ", b">>"]


def rand_walk(rng):
    kind = rng.integers(0, 10)
    if kind == 0:
        return rng.choice([b"*", b"", b">", b"<<", b"chr1", b"**"])
    n = int(rng.integers(1, 9))
    if kind == 1:   # monotone inside species a
        start = int(rng.integers(1, 50))
        ids = list(range(start, min(61, start + n)))
        if rng.integers(0, 2):
            ids.reverse()
    elif kind == 2:  # repeats
        ids = [int(x) for x in rng.integers(3, 9, n)]
    elif kind == 3:  # species b
        ids = [int(x) for x in rng.integers(61, 101, n)]
    elif kind == 4:  # spans species -> U
        ids = [59, 60, 61]
    elif kind == 5:  # out of every range / zero padded / too long runs
        ids = [int(rng.choice([0, 105, 4000000000, 12345678901]))] + [int(x) for x in rng.integers(1, 60, n)]
    elif kind == 6:  # species c (no graph)
        ids = [int(x) for x in rng.integers(101, 105, n)]
    else:
        ids = [int(x) for x in rng.integers(1, 61, n)]
    out = b""
    for v in ids:
        sep = SEPS[int(rng.integers(0, 2))] if rng.random() < 0.9 else SEPS[int(rng.integers(0, len(SEPS)))]
        txt = str(v).encode()
        if rng.random() < 0.05:
            txt = b"000" + txt
        if rng.random() < 0.02:
            txt = b"9" * int(rng.integers(19, 25))   # > 18 digits: dropped (rcls.rs:244)
        out += sep + txt
    return out


def rand_line(rng, i, wild_start):
    r = rng.random()
    if r < 0.03:
        return rng.choice([b"", b"\r", b"@HD\tVN:1", b"@", b"\t", b"\t\t\t\t\t\t\t\t\t\t\t\t"])
    pick = lambda: INTS[int(rng.integers(0, len(INTS)))] if rng.random() < 0.25 else str(int(rng.integers(0, 200))).encode()
    rid = IDS[int(rng.integers(0, len(IDS)))]
    rid = rid % i if b"%d" in rid else rid
    # column 8 (path start): a start beyond the first node is a panic in the reference (profile.rs:854) and voids the
    # whole species here, so most cases keep it at 0/1 (every node is at least 1 long); `wild_start` cases do not
    start = pick() if wild_start else rng.choice([b"0", b"1", b"0", b"*", b"-1", b"+0", b"", b"00"], p=[0.45, 0.3, 0.1, 0.03, 0.03, 0.03, 0.03, 0.03])
    cols = [rid, pick(), pick(), pick(), rng.choice([b"+", b"-", b"*"]), rand_walk(rng), pick(), start, pick(), pick(), pick(),
            rng.choice([b"60", b"0", b"3", b"2", b"59", b"61", b"*", b"255", b"", b"x"]), b"tp:A:P", b"cs:Z::10"]
    ncol = len(cols) if rng.random() < 0.85 else int(rng.integers(1, len(cols) + 1))
    line = b"\t".join(cols[:ncol])
    if rng.random() < 0.05:
        line += b"\r"
    if rng.random() < 0.02:
        line += b"\t"
    return line


def fuzz_gaf(seed, n, wild_start=False):
    rng = np.random.default_rng(seed)
    lines = [rand_line(rng, i, wild_start) for i in range(n)]
    gaf = b"\n".join(lines)
    if rng.random() < 0.5:
        gaf += b"\n"
    return gaf


@pytest.mark.parametrize("seed,wild_start", [(s, False) for s in range(8)] + [(100, True), (101, True)])
def test_three_cpu_implementations_agree_on_hostile_lines(seed, wild_start):
    gaf = fuzz_gaf(1000 + seed, 700, wild_start)
    if not wild_start:   # the coverage comparison must really happen
        from common import run_cpu_oracle
        o = run_cpu_oracle(RANGES, GRAPHS, gaf)
        assert o.species_error(0) == 0 and o.species_error(1) == 0 and int(o.node_bases(0).sum()) > 0 and o.mixed_dropped > 0
    assert_cpu_matches_py(RANGES, GRAPHS, gaf)                     # Python oracle == C++ oracle
    assert_hostcheck_matches(RANGES, GRAPHS, gaf)                  # kernel core == C++ oracle (also without the stash)
    assert_hostcheck_matches(RANGES, GRAPHS, gaf, stage_lim=48)    # tiny window: most lines take the re-parse path


@pytest.mark.gpu
@pytest.mark.parametrize("seed,wild_start,long_mode", [(0, False, None), (1, False, "1"), (2, False, "0"), (100, True, None), (101, True, "1")])
def test_gpu_agrees_on_hostile_lines(seed, wild_start, long_mode, monkeypatch):
    """The same hostile lines through the CUDA path (both ingest kernels), in one chunk and split into three."""
    from gpu_common import gpu_vs_oracle
    if long_mode is not None:
        monkeypatch.setenv("PTX_LONG_MODE", long_mode)
    gaf = fuzz_gaf(1000 + seed, 700, wild_start)
    gpu_vs_oracle(RANGES, GRAPHS, gaf)
    gpu_vs_oracle(RANGES, GRAPHS, gaf, split=[len(gaf) // 3, 2 * len(gaf) // 3])


# ---- gaf_filter.rs:22-97 -------------------------------------------------------------------------------------------
IDENT = [b"id:f:0.99", b"id:f:1", b"id:f:0.5", b"id:f:.75", b"id:f:1e-1", b"id:f:9.5E-1", b"id:f:-0.1", b"id:f:+0.25", b"id:f:abc", b"id:f:",
         b"id:f:0.123456789012345", b"0.9", b"x:y:0.8:0.7", b"id:f:inf", b"id:f:nan", b"id:f:1.", b"id:f:00.5", b"id:f:5e0"]
I32 = [b"0", b"10", b"21", b"20", b"60", b"1500", b"3000", b"-5", b"+7", b"2147483647", b"2147483648", b"-2147483648", b"x", b"", b"1.0", b" 3"]


def fuzz_filter_gaf(seed, n):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n):
        rid = b"q%d" % int(rng.integers(0, n // 3 + 1))          # several alignments per read
        p = lambda: I32[int(rng.integers(0, len(I32)))] if rng.random() < 0.15 else str(int(rng.integers(0, 4000))).encode()
        cols = [rid, p(), p(), p(), b"+", b">1>2", p(), p(), p(), p(), p(), p(), b"NM:i:1", b"AS:f:3", b"dv:f:0.01", IDENT[int(rng.integers(0, len(IDENT)))]]
        k = len(cols) if rng.random() < 0.9 else int(rng.integers(10, 17))
        if k > len(cols):
            cols = cols + [b"zz:Z:extra"]
        line = b"\t".join(cols[:k])
        if rng.random() < 0.05:
            line = b" " + line + b" "                           # the reference trims the line (gaf_filter.rs:23)
        if rng.random() < 0.05:
            line += b"\r"
        lines.append(line)
    return b"\n".join(lines) + b"\n"


@pytest.mark.parametrize("seed", range(6))
def test_filter_oracles_agree_on_hostile_lines(seed):
    from common import ocpu, opy
    gaf = fuzz_filter_gaf(50 + seed, 600)
    a = opy.filter_max_alignment(gaf)
    b = ocpu.filter_gaf(gaf)
    assert a == b
    assert 0 < len(a) < 600
