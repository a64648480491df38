"""Offline fuzz campaign over the three CPU implementations of the hot path (no GPU): the naive Python restatement, the C++
restatement and the exact code the kernels run (ptx_core.cuh / ptx_fast.cuh through tests/hostcheck.cpp, with the normal window, a
48-byte window that forces the re-parse path, and without the walk stash).  The committed tests run a dozen seeds of the same
generators; this runs thousands.

    python tools/fuzz_campaign.py FIRST_SEED LAST_SEED          # e.g. four of these side by side on disjoint seed ranges

Round 2: seeds 1000..6999 of the hostile-line generator in both modes (12,000 files, 7.2 M lines) and 1,000 synthetic data sets
x {short reads, short reads with duplicated ids, long reads}: 0 mismatches."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import test_fuzz_dialect as F  # noqa: E402
from common import NASTY, NASTY_DUP, assert_cpu_matches_py, dataset_graphs, synth  # noqa: E402
from test_core_host import assert_hostcheck_matches  # noqa: E402


def main() -> int:
    a, b = int(sys.argv[1]), int(sys.argv[2])
    t0, bad = time.time(), 0
    for seed in range(a, b):
        for wild in (False, True):
            gaf = F.fuzz_gaf(170000 + seed, 600, wild)
            try:
                assert_cpu_matches_py(F.RANGES, F.GRAPHS, gaf)
                assert_hostcheck_matches(F.RANGES, F.GRAPHS, gaf)
                assert_hostcheck_matches(F.RANGES, F.GRAPHS, gaf, stage_lim=48)
            except OverflowError:
                pass  # a species' sum of 19-digit read lengths left int64: the Python restatement's exact integers cannot be compared
            except AssertionError as e:
                bad += 1
                path = f"/tmp/fuzz_bad_{seed}_{int(wild)}.gaf"
                open(path, "wb").write(gaf)
                print("MISMATCH hostile lines, seed", seed, wild, "->", path, str(e)[:300], flush=True)
    print(f"hostile lines: seeds {a}..{b - 1} x 2 modes in {time.time() - t0:.0f} s, {bad} mismatches", flush=True)
    rng = np.random.default_rng(a)
    for seed in range(a, a + (b - a) // 6):
        nsp = int(rng.integers(1, 5))
        nodes = [int(x) for x in rng.integers(200, 6000, nsp)]
        haps = [int(x) for x in rng.integers(1, 7, nsp)]
        ds = synth.Dataset(9000 + seed, nodes, haps)
        graphs = dataset_graphs(ds)
        for params, n in ((NASTY, 3000), (NASTY_DUP, 3000), (synth.GafParams(long_reads=True, id_pair_suffix=False), 150)):
            gaf = ds.gaf(seed, 0, n, params)
            try:
                assert_hostcheck_matches(ds.ranges(), graphs, gaf)
                if n <= 150 or seed % 4 == 0:
                    assert_cpu_matches_py(ds.ranges(), graphs, gaf)
            except AssertionError as e:
                bad += 1
                print("MISMATCH synthetic data set, seed", seed, nodes, haps, params, str(e)[:300], flush=True)
    print(f"all done in {time.time() - t0:.0f} s, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
