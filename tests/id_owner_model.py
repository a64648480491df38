"""Host-side MODEL of the cross-rank read-id protocol that ptx_finalize runs on the GPUs (test infrastructure: used by
tests/test_sharding_gloo.py to check the protocol itself over gloo against the single-process oracle; not part of the product)."""
from __future__ import annotations


# ---- read-id groups across ranks: the host-side statement of the protocol ptx_finalize runs on the GPUs -------------
# (profile.rs:369-378 uniqueness over all non-U rows, :406-437 a group is kept only if all its eligible rows share
# one species - both are keyed by read id, which ignores shard boundaries.)
STATE_NONE = -1   # id seen, no coverage-eligible row yet
STATE_MIXED = -2  # eligible rows of several species


def merge_state(have: int, new: int) -> int:
    """State algebra of one id: NONE is the identity, equal species stay, different species -> MIXED (absorbing)."""
    if have == STATE_NONE:
        return new
    if new == STATE_NONE or have == new:
        return have
    return STATE_MIXED


class IdOwnerSet:
    """One rank's part of the distributed id set.  An id is kept only by the rank owning its hash; rows whose id
    another rank owns are queued in that rank's outbox as (hash, state) and merged there after the exchange."""

    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world
        self.table = {}                      # hash -> state, ids this rank owns
        self.outbox = [[] for _ in range(world)]
        self.repeat = False

    def _insert(self, h: int, state: int):
        if h in self.table:
            self.repeat = True
            self.table[h] = merge_state(self.table[h], state)
        else:
            self.table[h] = state

    def add_row(self, h: int, species: int, eligible: bool):
        """One non-U GAF row of this rank's read batch (k_apply<CLASSIFY>)."""
        state = species if eligible else STATE_NONE
        owner = h % self.world
        if owner == self.rank:
            self._insert(h, state)
        else:
            self.outbox[owner].append((h, state))

    def merge_inbox(self, entries):
        """Entries received from the other ranks (k_ds_merge_boxes)."""
        for h, state in entries:
            self._insert(h, state)

    def mixed_ids(self):
        return [h for h, s in self.table.items() if s == STATE_MIXED]
