"""`python -m pantax_b200 profile --db DB --gaf gfa_mapped.gaf --wd OUT [--long-read] [...]`: the profiling stage of the reference
(profile.rs:3325-3436 `profile`) from the aligner's GAF to the three tables, as two processes over the same files the reference uses:

  1. pantax-gpu-profile (C++ over the C ABI, GPU): reads_classification.tsv, species_abundance.txt, strain_inputs/, strain_graphs/
  2. strain_tail (host solver): strain_abundance.txt, ori_strain_abundance.txt

Options not listed here are handed to pantax-gpu-profile unchanged (`--ds`, `-a`, `--smode`, `-R`, `--chunk-mb`, ...).
`--long-read` first replaces the GAF by its best alignment per read (alignment.rs:171, `--filter-gaf`)."""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from typing import List, Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "pantax-gpu-profile")


def plan(argv: Sequence[str]) -> List[List[str]]:
    """The commands `profile` runs, in order (split out so that the plumbing can be checked without a GPU)."""
    ap = argparse.ArgumentParser(prog="python -m pantax_b200 profile")
    ap.add_argument("--db", "-d", required=True)
    ap.add_argument("--gaf", required=True)
    ap.add_argument("--wd", "-T", default=".")
    ap.add_argument("--long-read", action="store_true")
    ap.add_argument("--filter", action="store_true", help="with --long-read: run the best-alignment pre-filter on the GAF first (alignment.rs:171)")
    ap.add_argument("--species-only", action="store_true")
    ap.add_argument("--fr", type=float, default=None)
    ap.add_argument("--fc", type=float, default=None)
    ap.add_argument("--sr", type=float, default=None)
    ap.add_argument("--sd", type=float, default=None)
    ap.add_argument("--shift", action="store_true")
    ap.add_argument("--min_cov", type=float, default=None)
    ap.add_argument("--min_depth", type=float, default=None)
    ap.add_argument("--sample", type=int, default=None)
    a, rest = ap.parse_known_args(list(argv))
    cmds: List[List[str]] = []
    gaf = a.gaf
    if a.long_read and a.filter:
        if gaf == "-":
            raise SystemExit("--filter needs the GAF as a file")
        cmds.append([BIN, "--filter-gaf", gaf])
        d, base = os.path.split(gaf)
        stem = base[: base.rfind(".")] if "." in base[1:] else base
        gaf = os.path.join(d, stem + "_filtered.gaf")
    drv = [BIN, "--db", a.db, "--gaf", gaf, "--wd", a.wd, "--species"] + ([] if a.species_only else ["--strain"])
    tail = [sys.executable, "-m", "pantax_b200.strain_tail", "--db", a.db, "--wd", a.wd]
    if a.long_read:
        drv.append("--long-read")
        tail.append("--long-read")
    if a.shift:
        drv.append("--shift")
        tail.append("--shift")
    if a.fr is not None:
        drv += ["--fr", repr(a.fr)]
        tail += ["--fr", repr(a.fr)]
    if a.min_depth is not None:
        drv += ["--min-depth", repr(a.min_depth)]
        tail += ["--min_depth", repr(a.min_depth)]
    for name, v in (("--fc", a.fc), ("--sr", a.sr), ("--sd", a.sd), ("--min_cov", a.min_cov), ("--sample", a.sample)):
        if v is not None:
            tail += [name, repr(v)]
    cmds.append(drv + rest)
    if not a.species_only:
        cmds.append(tail)
    return cmds


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] != "profile":
        print(__doc__)
        return 0 if argv and argv[0] in ("-h", "--help") else 2
    if not os.path.exists(BIN):
        raise SystemExit(f"{BIN} not built; run `python -m pantax_b200.build` (there is no CPU fallback)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.dirname(HERE) + os.pathsep + env.get("PYTHONPATH", "")
    for cmd in plan(argv[1:]):
        rc = subprocess.call(cmd, env=env)
        if rc != 0:
            return rc
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
