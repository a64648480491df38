// The reference's own numbering of the unique trios (host only, plain C++).
//
// profile.rs:659-685 collects the canonical 3-windows of all paths in an FxHashSet<(usize, usize, usize)> - one
// `extend` per haplotype, haplotypes in BTreeMap (name) order - and numbers them in the order `into_iter()` yields them;
// the unique ones (count == 1, :707-716) keep that relative order.  The order is visible in exactly one place: the f64 sums
// of zscore_filter / frequencies_mean (profile.rs:1037-1041, 1146) run over a haplotype's trio abundances in index order.
// The device numbers trios by (owner haplotype, window position); this file reproduces the reference's permutation so
// that the host tail can add the same numbers in the same order.
//
// What is emulated (neither crate is under /root/reference; both restated from their published sources):
//  * fxhash 0.2.1 (Cargo.lock:1130) on a 64-bit target: FxHasher64, hash = (rotl(hash, 5) ^ word) * 0x517cc1b727220a95,
//    one word per tuple field (usize::hash -> write_usize), initial state 0.
//  * std::collections::HashSet = hashbrown's RawTable as vendored by the standard library since Rust 1.72 (hashbrown
//    >= 0.14; Cargo.lock pins polars 0.46, which needs a newer toolchain than that), x86-64 SSE2 flavour:
//      - buckets are a power of two: 4 below 4 elements, 8 below 8, else next_power_of_two(cap * 8 / 7);
//        usable capacity = buckets - 1 below 8 buckets, else buckets / 8 * 7;
//      - `extend` reserves size_hint when the set is empty, (size_hint + 1) / 2 otherwise, then inserts one by one;
//      - every insert first makes room for one element (reserve(1) precedes the lookup, so a table that is exactly full
//        grows even if the key turns out to be present), then probes: start = hash & mask, groups of 16 control bytes,
//        triangular steps (16, 32, ...); the key is searched group by group until a group with an empty byte; a new key
//        goes to the first empty byte of the first group on its probe sequence that has one (tables smaller than a group
//        wrap around, which is what the trailing control bytes + fix_insert_slot amount to);
//      - growing allocates max(items + additional, capacity + 1) and re-inserts the old buckets in ascending index order;
//      - nothing is ever removed, so there are no tombstones and no in-place rehash;
//      - iteration (`into_iter`) visits buckets in ascending index order.
//    On aarch64 (NEON, group width 8) or with a pre-1.72 toolchain the reference itself numbers the trios differently.
// Nothing in the reference's tests pins any of this: parity of the ORDER is unpinned; the set of trios is not affected.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace ptx_fx {

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fx_hash3(uint64_t a, uint64_t b, uint64_t c) {
    const uint64_t K = 0x517cc1b727220a95ULL;
    uint64_t h = a * K;  // (rotl(0, 5) ^ a) * K
    h = (rotl64(h, 5) ^ b) * K;
    h = (rotl64(h, 5) ^ c) * K;
    return h;
}

struct TrioSet {
    static constexpr uint64_t WIDTH = 16;  // SSE2 group
    uint64_t buckets = 0, items = 0, growth_left = 0;
    std::vector<uint8_t> full;
    std::vector<uint32_t> key;    // 3 per bucket
    std::vector<uint32_t> count;  // occurrences with multiplicity (profile.rs:689-702), saturating

    static uint64_t capacity_to_buckets(uint64_t cap) {
        if (cap < 8) return cap < 4 ? 4 : 8;
        uint64_t adj = cap * 8 / 7, b = 1;
        while (b < adj) b <<= 1;
        return b;
    }
    static uint64_t buckets_to_capacity(uint64_t b) {
        if (b == 0) return 0;
        return b <= 8 ? b - 1 : b / 8 * 7;
    }
    uint64_t insert_slot(uint64_t hash) const {
        const uint64_t mask = buckets - 1;
        uint64_t pos = hash & mask, stride = 0;
        for (;;) {
            for (uint64_t k = 0; k < WIDTH; ++k) {
                uint64_t i = (pos + k) & mask;
                if (!full[i]) return i;
            }
            stride += WIDTH;
            pos = (pos + stride) & mask;
        }
    }
    void resize(uint64_t cap) {
        uint64_t nb = capacity_to_buckets(cap);
        TrioSet t;
        t.buckets = nb;
        t.full.assign(nb, 0);
        t.key.resize(3 * nb);
        t.count.resize(nb);
        for (uint64_t i = 0; i < buckets; ++i) {
            if (!full[i]) continue;
            uint64_t j = t.insert_slot(fx_hash3(key[3 * i], key[3 * i + 1], key[3 * i + 2]));
            t.full[j] = 1;
            memcpy(&t.key[3 * j], &key[3 * i], 12);
            t.count[j] = count[i];
        }
        t.items = items;
        t.growth_left = buckets_to_capacity(nb) - items;
        *this = std::move(t);
    }
    void reserve(uint64_t additional) {
        if (additional <= growth_left) return;
        resize(std::max(items + additional, buckets_to_capacity(buckets) + 1));
    }
    void insert(uint32_t a, uint32_t b, uint32_t c) {
        reserve(1);
        const uint64_t mask = buckets - 1, hash = fx_hash3(a, b, c);
        uint64_t pos = hash & mask, stride = 0, slot = ~0ULL;
        for (;;) {
            bool any_empty = false;
            for (uint64_t k = 0; k < WIDTH; ++k) {
                uint64_t i = (pos + k) & mask;
                if (full[i]) {
                    if (key[3 * i] == a && key[3 * i + 1] == b && key[3 * i + 2] == c) {
                        if (count[i] != 0xffffffffu) ++count[i];
                        return;
                    }
                } else {
                    any_empty = true;
                    if (slot == ~0ULL) slot = i;
                }
            }
            if (any_empty) break;
            stride += WIDTH;
            pos = (pos + stride) & mask;
        }
        full[slot] = 1;
        key[3 * slot] = a, key[3 * slot + 1] = b, key[3 * slot + 2] = c;
        count[slot] = 1;
        ++items;
        --growth_left;
    }
    void extend_reserve(uint64_t n) { reserve(items == 0 ? n : (n + 1) / 2); }
};

struct Key3 {
    uint32_t a, b, c;
    uint64_t idx;
    bool operator<(const Key3& o) const { return a != o.a ? a < o.a : b != o.b ? b < o.b : c < o.c; }
};

// order[i] = row of the library's trio table (keys3, n_trios rows of canonical local ids) that the reference numbers i.
// `all_out` (optional): every distinct trio in the reference's iteration order, 3 ids each.  Returns 0, or -1 when the paths and
// the table do not describe the same set of unique trios (or an id does not fit 32 bits).
static inline int trio_ref_order(const uint64_t* path_off, const uint64_t* path_nodes, int64_t n_paths, const uint64_t* keys3,
                                 int64_t n_trios, uint64_t* order, std::vector<uint32_t>* all_out = nullptr) {
    TrioSet set;
    for (int64_t h = 0; h < n_paths; ++h) {
        const uint64_t b = path_off[h], e = path_off[h + 1];
        const uint64_t w = e - b >= 3 ? e - b - 2 : 0;
        set.extend_reserve(w);
        for (uint64_t i = 0; i < w; ++i) {
            uint64_t x = path_nodes[b + i], y = path_nodes[b + i + 1], z = path_nodes[b + i + 2];
            if ((x | y | z) >> 32) return -1;
            if (x > z) std::swap(x, z);
            set.insert((uint32_t)x, (uint32_t)y, (uint32_t)z);
        }
    }
    std::vector<Key3> tab((size_t)n_trios);
    for (int64_t t = 0; t < n_trios; ++t) {
        if ((keys3[3 * t] | keys3[3 * t + 1] | keys3[3 * t + 2]) >> 32) return -1;
        tab[(size_t)t] = Key3{(uint32_t)keys3[3 * t], (uint32_t)keys3[3 * t + 1], (uint32_t)keys3[3 * t + 2], (uint64_t)t};
    }
    std::sort(tab.begin(), tab.end());
    int64_t n = 0;
    if (all_out) all_out->clear();
    for (uint64_t i = 0; i < set.buckets; ++i) {
        if (!set.full[i]) continue;
        if (all_out) all_out->insert(all_out->end(), &set.key[3 * i], &set.key[3 * i] + 3);
        if (set.count[i] != 1) continue;
        Key3 k{set.key[3 * i], set.key[3 * i + 1], set.key[3 * i + 2], 0};
        auto it = std::lower_bound(tab.begin(), tab.end(), k);
        if (it == tab.end() || k < *it || n >= n_trios) return -1;
        order[n++] = it->idx;
    }
    return n == n_trios ? 0 : -1;
}

}  // namespace ptx_fx
