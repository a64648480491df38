"""Read-batch sharding of a GAF byte stream across ranks.

Rank k of P takes the byte range [k*B/P, (k+1)*B/P) snapped FORWARD to the next line start,
so every line belongs to exactly one rank and no rank needs another's bytes.  Integer sums
and bit-ORs are associative and commutative, hence the reduced result is bit-identical for
any P.  (Host-side logic only; the reductions are in ptx_finalize.)
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def shard_bounds(n_bytes: int, world: int, next_line_start: Callable[[int], int]) -> List[Tuple[int, int]]:
    """`next_line_start(pos)` returns the offset of the first line start >= pos (pos itself if
    pos == 0 or the byte before pos is a newline), or n_bytes if there is none."""
    cuts = [0]
    for k in range(1, world):
        cuts.append(min(n_bytes, max(cuts[-1], next_line_start(n_bytes * k // world))))
    cuts.append(n_bytes)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def shard_bounds_bytes(data: bytes, world: int) -> List[Tuple[int, int]]:
    n = len(data)

    def nls(pos: int) -> int:
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        if data[pos - 1:pos] == b"\n":
            return pos
        j = data.find(b"\n", pos)
        return n if j < 0 else j + 1

    return shard_bounds(n, world, nls)
