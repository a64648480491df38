"""TEST INFRASTRUCTURE (oracle) - literal restatement of the hash set behind the reference's trio numbering.

The reference numbers trios in the iteration order of `FxHashSet<(usize, usize, usize)>` (profile.rs:659-685).  Neither crate
is under /root/reference; this file restates their published algorithms *structurally* - control-byte array with its
trailing mirror bytes, 7-bit tags, group loads at unaligned positions, `fix_insert_slot` for tables smaller than a group -
so that it can check the product's simplified emulation (`pantax_b200/csrc/ptx_fxorder.h`, circular 16-slot scans, no tags).
Parity unpinned: nothing in the reference's tests fixes this order, and no Rust toolchain exists here to run it.

  fxhash 0.2.1 (Cargo.lock:1130)  FxHasher64: hash = (rotl(hash, 5) ^ word) * 0x517cc1b727220a95 per usize field
  std HashSet = hashbrown >= 0.14 (as vendored by std since Rust 1.72), SSE2 Group::WIDTH = 16:
      capacity_to_buckets, bucket_mask_to_capacity, reserve_rehash -> resize (re-insert in ascending bucket order),
      RawTable::find_or_find_insert_slot (reserve(1) first), HashMap::extend's reserve rule, RawIter order.
Only tests/ may import this module.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

MASK64 = (1 << 64) - 1
SEED64 = 0x517CC1B727220A95
EMPTY = 0xFF
WIDTH = 16

Trio = Tuple[int, int, int]


def fx_hash_words(words: Iterable[int]) -> int:
    h = 0
    for w in words:
        h = ((((h << 5) | (h >> 59)) & MASK64) ^ (w & MASK64)) * SEED64 & MASK64
    return h


def h2(hash_: int) -> int:
    return (hash_ >> 57) & 0x7F  # top 7 bits


def capacity_to_buckets(cap: int) -> int:
    assert cap != 0
    if cap < 8:
        return 4 if cap < 4 else 8
    adjusted = cap * 8 // 7
    b = 1
    while b < adjusted:
        b <<= 1
    return b


def bucket_mask_to_capacity(mask: int) -> int:
    return mask if mask < 8 else (mask + 1) // 8 * 7


class RawTable:
    """hashbrown::raw::RawTableInner for a set of hashable keys (no removal)."""

    def __init__(self) -> None:
        # the static empty singleton: one group of EMPTY bytes, bucket_mask 0
        self.bucket_mask = 0
        self.ctrl = [EMPTY] * WIDTH
        self.data: List = [None]
        self.items = 0
        self.growth_left = 0
        self.singleton = True

    @property
    def buckets(self) -> int:
        return self.bucket_mask + 1

    @classmethod
    def with_capacity(cls, cap: int) -> "RawTable":
        t = cls()
        if cap == 0:
            return t
        b = capacity_to_buckets(cap)
        t.bucket_mask = b - 1
        t.ctrl = [EMPTY] * (b + WIDTH)
        t.data = [None] * b
        t.growth_left = bucket_mask_to_capacity(t.bucket_mask)
        t.singleton = False
        return t

    def group(self, pos: int) -> List[int]:
        return self.ctrl[pos:pos + WIDTH]

    def set_ctrl(self, index: int, byte: int) -> None:
        index2 = ((index - WIDTH) & self.bucket_mask) + WIDTH
        self.ctrl[index] = byte
        self.ctrl[index2] = byte

    def is_bucket_full(self, index: int) -> bool:
        return self.ctrl[index] & 0x80 == 0

    def fix_insert_slot(self, index: int) -> int:
        if self.is_bucket_full(index):
            assert self.bucket_mask < WIDTH
            g = self.group(0)
            index = next(i for i, c in enumerate(g) if c & 0x80)
        return index

    def find_insert_slot(self, hash_: int) -> int:
        pos, stride = hash_ & self.bucket_mask, 0
        while True:
            g = self.group(pos)
            for bit, c in enumerate(g):
                if c & 0x80:
                    return self.fix_insert_slot((pos + bit) & self.bucket_mask)
            stride += WIDTH
            pos = (pos + stride) & self.bucket_mask

    def resize(self, capacity: int, hasher) -> None:
        new = RawTable.with_capacity(capacity)
        for i in range(self.buckets):  # full_buckets_indices(): ascending
            if self.singleton or not self.is_bucket_full(i):
                continue
            hv = hasher(self.data[i])
            j = new.find_insert_slot(hv)
            new.set_ctrl(j, h2(hv))
            new.data[j] = self.data[i]
        new.items = self.items
        new.growth_left -= self.items
        self.__dict__.update(new.__dict__)

    def reserve(self, additional: int, hasher) -> None:
        if additional > self.growth_left:
            new_items = self.items + additional
            full_capacity = 0 if self.singleton else bucket_mask_to_capacity(self.bucket_mask)
            assert not new_items <= full_capacity // 2  # would be rehash_in_place: impossible without tombstones
            self.resize(max(new_items, full_capacity + 1), hasher)

    def find_or_find_insert_slot(self, hash_: int, key, hasher):
        self.reserve(1, hasher)
        tag = h2(hash_)
        insert_slot = None
        pos, stride = hash_ & self.bucket_mask, 0
        while True:
            g = self.group(pos)
            for bit, c in enumerate(g):
                if c == tag:
                    index = (pos + bit) & self.bucket_mask
                    if self.data[index] == key:
                        return True, index
            if insert_slot is None:
                for bit, c in enumerate(g):
                    if c & 0x80:
                        insert_slot = (pos + bit) & self.bucket_mask
                        break
            if any(c == EMPTY for c in g):
                return False, self.fix_insert_slot(insert_slot)
            stride += WIDTH
            pos = (pos + stride) & self.bucket_mask

    def insert_in_slot(self, hash_: int, slot: int, key) -> None:
        old = self.ctrl[slot]
        self.growth_left -= 1 if old == EMPTY else 0
        self.set_ctrl(slot, h2(hash_))
        self.data[slot] = key
        self.items += 1


class FxHashSet:
    def __init__(self, hasher=lambda k: fx_hash_words(k)) -> None:
        self.t = RawTable()
        self.hasher = hasher

    def __len__(self) -> int:
        return self.t.items

    def insert(self, key) -> bool:
        hv = self.hasher(key)
        found, slot = self.t.find_or_find_insert_slot(hv, key, self.hasher)
        if found:
            return False
        self.t.insert_in_slot(hv, slot, key)
        return True

    def extend(self, keys: Sequence) -> None:
        """hashbrown HashMap::extend: reserve the whole size hint when empty, half of it otherwise."""
        n = len(keys)
        self.t.reserve(n if self.t.items == 0 else (n + 1) // 2, self.hasher)
        for k in keys:
            self.insert(k)

    def into_iter(self) -> List:
        if self.t.singleton:
            return []
        return [self.t.data[i] for i in range(self.t.buckets) if self.t.is_bucket_full(i)]


def canonical_windows(path: Sequence[int]) -> List[Trio]:
    """profile.rs:671-680."""
    return [((c, b, a) if a > c else (a, b, c)) for a, b, c in zip(path, path[1:], path[2:])]


def reference_trio_numbering(paths: Sequence[Sequence[int]]) -> Tuple[List[Trio], List[Trio]]:
    """profile.rs:659-716 for paths in BTreeMap (name) order: (all distinct trios in index order, the unique ones in their order)."""
    s = FxHashSet()
    per_hap = []
    for p in paths:
        w = canonical_windows(list(p))
        s.extend(w)
        per_hap.append(w)
    trio_nodes = s.into_iter()
    index = {t: i for i, t in enumerate(trio_nodes)}
    count = [0] * len(trio_nodes)
    for w in per_hap:
        for t in w:
            count[index[t]] += 1
    return trio_nodes, [t for i, t in enumerate(trio_nodes) if count[i] == 1]
