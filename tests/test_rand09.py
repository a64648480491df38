"""rand 0.9.2 `StdRng::seed_from_u64` + `choose_multiple` restated for profile.rs:1287-1295 (pantax_b200/rand09.py).
Known answers: the ChaCha block function against RFC 7539 section 2.3.2 (20 rounds), the all-zero-key keystreams of ChaCha20 /
ChaCha12 / ChaCha8 (eSTREAM vectors), and rand_chacha's own `test_chacha_construction` value (seed layout, counter placement).
The seeding by PCG32 and the index samplers have no published vector: structure and properties only (unpinned)."""
import numpy as np
import pytest

from pantax_b200 import rand09 as r
from pantax_b200 import strain_tail


def test_chacha_block_rfc7539():
    key = np.frombuffer(bytes(range(32)), dtype="<u4")
    out = r.chacha_blocks(key, np.array([[1, 0x09000000, 0x4A000000, 0]], dtype=np.uint32), 10)[0]
    want = "e4e7f110 15593bd1 1fdd0f50 c47120a3 c7f4d1c7 0368c033 9aaa2204 4e6cd4c3 466482d2 09aa9f07 05d7c214 a2028bd9 d19c12b5 b94e16de e883d0cb 4e3c50a2"
    assert " ".join(f"{x:08x}" for x in out) == want


def test_zero_key_keystreams():
    z, t = np.zeros(8, dtype=np.uint32), np.zeros((1, 4), dtype=np.uint32)
    ks = lambda dr: r.chacha_blocks(z, t, dr)[0].astype("<u4").tobytes().hex()
    assert ks(10).startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    assert ks(6) == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                     "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    assert ks(4).startswith("3e00ef2f895f40d67f5bb8e81f09a5a1")


def test_rand_chacha_seed_layout():
    """rand_chacha `test_chacha_construction`: ChaCha20Rng::from_seed([0,0,0,0,0,0,0,0,1,0,..,2,0,..,3,0,..]).next_u32() == 137206642."""
    seed = bytes([0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0])
    key = np.frombuffer(seed, dtype="<u4")
    assert int(r.chacha_blocks(key, np.zeros((1, 4), dtype=np.uint32), 10)[0][0]) == 137206642


def test_block_counter_is_64_bit_words_12_13_and_buffer_is_sequential():
    g = r.StdRng.seed_from_u64(7)
    words = [g.next_u32() for _ in range(16 * 5)]
    for b in range(5):
        blk = r.chacha_blocks(g.key, np.array([[b, 0, 0, 0]], dtype=np.uint32), 6)[0]
        assert words[16 * b:16 * b + 16] == [int(x) for x in blk]
    # refill continues the counter
    g2 = r.StdRng.seed_from_u64(7)
    g2.BATCH_BLOCKS = 2
    assert [g2.next_u32() for _ in range(80)] == words


def test_pcg32_seed_expansion():
    s = r.pcg32_seed_bytes(0)
    assert len(s) == 32 and len(set(s)) > 16
    # by hand, first word of seed 0: state = INC; xorshift / rotate
    st = 11634580027462260723
    xs = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
    rot = st >> 59
    x = ((xs >> rot) | (xs << (32 - rot))) & 0xFFFFFFFF
    assert s[:4] == x.to_bytes(4, "little")
    assert r.pcg32_seed_bytes(42) != r.pcg32_seed_bytes(43)


class Words:
    """A scripted u32 source with StdRng's sampling methods."""

    def __init__(self, words):
        self.w, self.i = list(words), 0

    def next_u32(self):
        v = self.w[self.i]
        self.i += 1
        return v

    random_range_u32 = r.StdRng.random_range_u32


def test_random_range_canon_biased():
    # range 10: word 0 -> 0 with low half 0 (no second draw)
    g = Words([0])
    assert g.random_range_u32(5, 15) == 5 and g.i == 1
    # low half above range.wrapping_neg() triggers ONE more draw; a carry bumps the result
    rng_ = 10
    w = ((2 ** 31 - 1) * pow(5, -1, 2 ** 31)) % 2 ** 31  # 10 w = -2 (mod 2^32): the low half is 2^32 - 2 > 2^32 - 10
    assert (w * rng_) & 0xFFFFFFFF == 2 ** 32 - 2
    base, lo = (w * rng_) >> 32, (w * rng_) & 0xFFFFFFFF
    g = Words([w, 0xFFFFFFFF])
    assert g.random_range_u32(0, 10) == base + (1 if lo + ((0xFFFFFFFF * rng_) >> 32) > 0xFFFFFFFF else 0) and g.i == 2
    g = Words([w, 0])
    assert g.random_range_u32(0, 10) == base and g.i == 2
    # full range: one raw word
    g = Words([123456])
    assert g.random_range_u32(0, 2 ** 32) == 123456


def test_algorithm_choice_follows_index_sample():
    calls = []
    orig = (r._sample_inplace, r._sample_floyd, r._sample_rejection)
    try:
        r._sample_inplace = lambda g, l, a: calls.append("inplace") or np.zeros(a, dtype=np.uint32)
        r._sample_floyd = lambda g, l, a: calls.append("floyd") or np.zeros(a, dtype=np.uint32)
        r._sample_rejection = lambda g, l, a: calls.append("rejection") or np.zeros(a, dtype=np.uint32)
        for length, amount in ((600_000, 500_000), (18_000_000, 500_000), (18_400_000, 500_000), (100_000, 500), (140_000, 500),
                               (600_000, 500), (1000, 5), (1000, 100), (20_000, 100)):
            r.index_sample(None, length, amount)
    finally:
        r._sample_inplace, r._sample_floyd, r._sample_rejection = orig
    assert calls == ["inplace", "inplace", "rejection", "inplace", "rejection", "rejection", "floyd", "inplace", "floyd"]


@pytest.mark.parametrize("length,amount", [(2000, 500), (200_000, 500), (50, 50), (40, 7), (3000, 100)])
def test_samples_are_distinct_in_range_and_deterministic(length, amount):
    vals = np.arange(length) * 7 + 3
    a = r.choose_multiple_sorted(vals, amount, 42)
    b = r.choose_multiple_sorted(vals, amount, 42)
    c = r.choose_multiple_sorted(vals, amount, 43)
    assert np.array_equal(a, b) and len(a) == amount and len(np.unique(a)) == amount
    assert np.all(np.diff(a) > 0) and np.isin(a, vals).all()
    if amount < length:
        assert not np.array_equal(a, c)


def test_inplace_is_a_partial_fisher_yates_of_the_word_stream():
    g = r.StdRng.seed_from_u64(42)
    h = r.StdRng.seed_from_u64(42)
    got = r._sample_inplace(g, 1000, 300)
    idx = list(range(1000))
    for i in range(300):
        j = h.random_range_u32(i, 1000)
        idx[i], idx[j] = idx[j], idx[i]
    assert got.tolist() == idx[:300]


def test_strain_tail_uses_it():
    v = np.arange(0, 3000, 3)
    assert np.array_equal(strain_tail.sample_sorted(v, 500, 42), r.choose_multiple_sorted(v, 500, 42))
    # uniformity sanity: every tenth of the range gets its share
    s = r.choose_multiple_sorted(np.arange(100_000), 20_000, 42)
    hist = np.histogram(s, bins=10, range=(0, 100_000))[0]
    assert hist.min() > 1800 and hist.max() < 2200
