"""`StdRng::seed_from_u64(seed)` + `slice.choose_multiple(&mut rng, amount)` of rand 0.9.2 (Cargo.lock) on the host.

The reference subsamples the covered nodes of a species before it builds the PAO rows (profile.rs:1287-1295, called with
seed 42 and `--sample` = 500,000 at :1397 and its five siblings); which nodes are drawn decides the ILP, so the draw is
reproduced rather than replaced.  rand is not under /root/reference; this restates its published algorithms:

  * rand_core `SeedableRng::seed_from_u64`: the 32-byte seed is eight outputs of a PCG32 stream
    (state = state * 6364136223846793005 + 11634580027462260723; xorshift 18/27, rotate by the top 5 bits), little endian.
  * `StdRng` = `ChaCha12Rng` (rand_chacha 0.9): key = seed, 64-bit block counter in words 12-13, stream id 0 in words 14-15,
    six double rounds; the block RNG buffers four consecutive blocks (64 words) and `next_u32` hands them out in order.
  * `index::sample(rng, length, amount)` (rand::seq::index): for amount >= 163 it runs `sample_inplace` when
    `length < C * amount` (C = 270 below 500,000 elements, 330/9 above, compared in f32) and `sample_rejection` otherwise;
    for amount < 163 Floyd's algorithm unless `amount > 11 && length < (C1 + C0 * amount) * amount` (-> in place).
  * `sample_inplace`: a partial Fisher-Yates over 0..length with `j = rng.random_range(i..length)` for i in 0..amount.
  * `random_range` on u32 (`UniformInt::sample_single_inclusive`, the default *biased* Canon variant of 0.9): one u32 draw,
    widening multiply by the range; if the low half exceeds `range.wrapping_neg()` a second draw's high half is added to it and
    a carry bumps the result.  `Uniform::sample` (used by the rejection sampler) is Lemire's method with
    `thresh = range.wrapping_neg() % range`.

The reference sorts the drawn elements (`sort_unstable`), so only the drawn SET matters.
Known answers: the ChaCha core is checked against the RFC 7539 block vector (20 rounds) in tests/test_rand09.py; the 12-round
stream, the seeding and the samplers have no vector in the reference - **unpinned**, like every third-party piece here.
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1


def pcg32_seed_bytes(state: int, n_words: int = 8) -> bytes:
    """rand_core::SeedableRng::seed_from_u64."""
    out = bytearray()
    for _ in range(n_words):
        state = (state * 6364136223846793005 + 11634580027462260723) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & M32
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & M32
        out += x.to_bytes(4, "little")
    return bytes(out)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def chacha_blocks(key_words: np.ndarray, tail4: np.ndarray, double_rounds: int) -> np.ndarray:
    """ChaCha block function for a batch: key_words u32[8], tail4 u32[n, 4] = state words 12..15 of each block -> u32[n, 16]."""
    n = tail4.shape[0]
    const = np.array([0x61707865, 0x3320646E, 0x79622D32, 0x6B206574], dtype=np.uint32)
    init = np.empty((16, n), dtype=np.uint32)
    init[0:4] = const[:, None]
    init[4:12] = np.asarray(key_words, dtype=np.uint32)[:, None]
    init[12:16] = np.asarray(tail4, dtype=np.uint32).T
    x = init.copy()

    def qr(a, b, c, d):
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(double_rounds):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        x += init
    return x.T.copy()


class StdRng:
    """rand 0.9 StdRng (ChaCha12, stream 0) as a u32 word source."""

    BATCH_BLOCKS = 4096

    def __init__(self, seed_bytes: bytes):
        assert len(seed_bytes) == 32
        self.key = np.frombuffer(seed_bytes, dtype="<u4").astype(np.uint32)
        self.counter = 0
        self.buf = np.zeros(0, dtype=np.uint32)
        self.pos = 0

    @classmethod
    def seed_from_u64(cls, seed: int) -> "StdRng":
        return cls(pcg32_seed_bytes(seed & M64))

    def _refill(self) -> None:
        c = np.arange(self.counter, self.counter + self.BATCH_BLOCKS, dtype=np.uint64)
        tail = np.zeros((self.BATCH_BLOCKS, 4), dtype=np.uint32)
        tail[:, 0] = (c & np.uint64(M32)).astype(np.uint32)
        tail[:, 1] = (c >> np.uint64(32)).astype(np.uint32)
        self.buf = chacha_blocks(self.key, tail, 6).reshape(-1)
        self.counter += self.BATCH_BLOCKS
        self.pos = 0

    def next_u32(self) -> int:
        if self.pos >= len(self.buf):
            self._refill()
        v = int(self.buf[self.pos])
        self.pos += 1
        return v

    def random_range_u32(self, low: int, high: int) -> int:
        """`rng.random_range(low..high)` for u32: UniformInt::sample_single -> sample_single_inclusive(low, high - 1)."""
        assert low < high
        rng_ = (high - 1 - low + 1) & M32
        if rng_ == 0:
            return self.next_u32()
        m = self.next_u32() * rng_
        result, lo_order = m >> 32, m & M32
        if lo_order > ((-rng_) & M32):
            new_hi = (self.next_u32() * rng_) >> 32
            if lo_order + new_hi > M32:
                result += 1
        return (low + result) & M32


def _sample_inplace(rng: StdRng, length: int, amount: int) -> np.ndarray:
    idx = np.arange(length, dtype=np.uint32)
    for i in range(amount):
        j = rng.random_range_u32(i, length)
        idx[i], idx[j] = idx[j], idx[i]
    return idx[:amount].copy()


def _sample_floyd(rng: StdRng, length: int, amount: int) -> np.ndarray:
    # rand 0.9 index::sample_floyd: for j in length - amount .. length: t = random_range(..=j); if t already drawn, push j
    indices = []
    for j in range(length - amount, length):
        t = rng.random_range_u32(0, j + 1)
        if t in indices:
            indices[indices.index(t)] = j
        indices.append(t)
    return np.array(indices, dtype=np.uint32)


def _sample_rejection(rng: StdRng, length: int, amount: int) -> np.ndarray:
    thresh = ((-length) & M32) % length  # Uniform::new(0, length): Lemire's threshold
    seen = set()
    out = []
    for _ in range(amount):
        while True:
            while True:
                m = rng.next_u32() * length
                if (m & M32) >= thresh:
                    pos = m >> 32
                    break
            if pos not in seen:
                seen.add(pos)
                break
        out.append(pos)
    return np.array(out, dtype=np.uint32)


def index_sample(rng: StdRng, length: int, amount: int) -> np.ndarray:
    """rand::seq::index::sample for length <= u32::MAX."""
    if amount > length:
        raise ValueError("`amount` of samples must be less than or equal to `length`")
    if length > M32:
        raise NotImplementedError("more than 2^32 - 1 elements")
    f32 = np.float32
    if amount < 163:
        C = ((f32(1.6), f32(8.0) / f32(45.0)), (f32(10.0), f32(70.0) / f32(9.0)))
        j = 1 if length >= 500_000 else 0
        amount_fp = f32(amount)
        m4 = C[0][j] * amount_fp
        if amount > 11 and f32(length) < (C[1][j] + m4) * amount_fp:
            return _sample_inplace(rng, length, amount)
        return _sample_floyd(rng, length, amount)
    C = (f32(270.0), f32(330.0) / f32(9.0))
    j = 1 if length >= 500_000 else 0
    if f32(length) < C[j] * f32(amount):
        return _sample_inplace(rng, length, amount)
    return _sample_rejection(rng, length, amount)


def choose_multiple_sorted(values: np.ndarray, amount: int, seed: int) -> np.ndarray:
    """profile.rs:1287-1295 `sample_sorted`: choose_multiple clamps `amount` to the slice length; the result is sorted."""
    values = np.asarray(values)
    amount = min(int(amount), len(values))
    idx = index_sample(StdRng.seed_from_u64(seed), len(values), amount)
    return np.sort(values[idx.astype(np.int64)])
