// Per-record logic of the PanTax hot path, written once as host/device inline code.
//
// The CUDA kernels in ptx_kernels.cu instantiate it with a device "sink" (atomics on
// HBM-resident accumulators); tests/hostcheck.cpp instantiates the very same code
// with a plain-array sink so that the parsing / span arithmetic can be checked on a
// box without a GPU.  The host instantiation is a TEST harness only - the shipped
// library has no CPU path.
//
// Reference semantics implemented here (paths relative to /root/reference/pantax/src):
//   GAF columns 1,2,6,7,8,9,12, '*' = null, '@' comment lines     rcls.rs:119-146
//   digit runs of the walk, min/max, first matching species range  rcls.rs:237-258
//   per-read node spans, first-occurrence rule, covered bits,
//   3-window (trio) base sums                                      profile.rs:787-919
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PTX_HD __host__ __device__ __forceinline__
#else
#define PTX_HD inline
#endif

namespace ptx {

constexpr uint32_t LABEL_U = 0xFFFFFFFFu;
constexpr int64_t NULL_I64 = INT64_MIN;
constexpr uint32_t TT_EMPTY = 0xFFFFFFFFu;

// dup-set slot state (low 32 bits of the second word)
constexpr uint32_t DS_NONE = 0xFFFFFFFEu;   // id seen, no coverage-eligible row yet
constexpr uint32_t DS_MIXED = 0xFFFFFFFDu;  // eligible rows of >1 species (profile.rs:415-416)

enum Term : int { T_TAB = 0, T_EOL = 1, T_LIMIT = 2 };

struct IdHash {  // 96 bits of read-id hash; lo is never 0 (0/0 marks an empty slot)
    uint64_t lo;
    uint32_t hi;
};

PTX_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
PTX_HD uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

// Three-lane murmur3-style hash over 4-byte little-endian words of the id.
struct IdHasher {
    uint32_t h0 = 0x9747b28cu, h1 = 0x3c6ef372u, h2 = 0xa54ff53au;
    uint32_t word = 0, nbytes = 0;
    PTX_HD void mix(uint32_t k) {
        k *= 0xcc9e2d51u;
        k = rotl32(k, 15);
        k *= 0x1b873593u;
        h0 ^= k;
        h0 = rotl32(h0, 13) * 5u + 0xe6546b64u;
        uint32_t t = h0;  // rotate lanes so every word touches all three
        h0 = h1 + t;
        h1 = h2 ^ rotl32(t, 7);
        h2 = t;
    }
    PTX_HD void byte(uint8_t c) {
        word |= (uint32_t)c << ((nbytes & 3u) * 8u);
        ++nbytes;
        if ((nbytes & 3u) == 0) { mix(word); word = 0; }
    }
    PTX_HD IdHash finish() {
        if (nbytes & 3u) mix(word);
        h0 ^= nbytes; h1 ^= nbytes * 0x9e3779b9u; h2 ^= ~nbytes;
        h0 += h1; h0 += h2; h1 += h0; h2 += h0;
        h0 = fmix32(h0); h1 = fmix32(h1); h2 = fmix32(h2);
        h0 += h1; h0 += h2; h1 += h0; h2 += h0;
        IdHash r;
        r.lo = ((uint64_t)h1 << 32) | (uint64_t)(h0 | 1u);
        r.hi = h2;
        return r;
    }
};

PTX_HD uint32_t trio_hash(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = a * 0x9e3779b1u;
    h = rotl32(h, 13) ^ (b * 0x85ebca77u);
    h = rotl32(h, 15) ^ (c * 0xc2b2ae3du);
    return fmix32(h);
}

struct RangesView {
    const int64_t* start;      // [S] 1-based inclusive, file order (species_range.txt)
    const int64_t* end;        // [S]
    const int64_t* node_base;  // [S] offset of the species in the concatenated node arrays, -1 = no graph
    const uint32_t* order;     // [S] species indices sorted by start (used when disjoint)
    int S;
    int disjoint;
};

// rcls.rs:253-257: FIRST range in file order with min>=start && max<=end.
PTX_HD uint32_t classify(const RangesView& R, int64_t lo, int64_t hi) {
    if (R.disjoint) {
        // ranges are pairwise disjoint: at most one can contain `lo`; binary search by start
        int a = 0, b = R.S;  // last position with start <= lo
        while (a < b) {
            int m = (a + b) >> 1;
            if (R.start[R.order[m]] <= lo) a = m + 1; else b = m;
        }
        if (a == 0) return LABEL_U;
        uint32_t s = R.order[a - 1];
        return (hi <= R.end[s]) ? s : LABEL_U;
    }
    for (int s = 0; s < R.S; ++s)
        if (lo >= R.start[s] && hi <= R.end[s]) return (uint32_t)s;
    return LABEL_U;
}

struct RecParse {
    IdHash h;
    int64_t qlen, c7, c8, c9, mapq;  // NULL_I64 = null
    uint32_t path_pos, path_end;      // bytes [pos,end) of column 6
    uint32_t W;                       // walk nodes (digit runs of <= 18 digits)
    int64_t vmin, vmax;
    bool path_null;
    bool monotone;                    // strictly increasing or strictly decreasing ids => no repeats
};

// Is b[p] (== c, not a digit) a field/line terminator?  -1: ordinary byte.
PTX_HD int term_at(const uint8_t* b, uint32_t p, uint32_t lim, uint8_t c) {
    if (c == '\t') return T_TAB;
    if (c == '\n') return T_EOL;
    if (c == '\r') {
        if (p + 1 >= lim) return T_LIMIT;
        if (b[p + 1] == '\n') return T_EOL;
    }
    return -1;
}

// Advance to the end of the current field.  On T_TAB the tab is consumed.
PTX_HD int skip_field(const uint8_t* b, uint32_t& p, uint32_t lim) {
    for (;;) {
        if (p >= lim) return T_LIMIT;
        uint8_t c = b[p];
        if (c <= '\r') {  // '\t'=9 '\n'=10 '\r'=13
            int t = term_at(b, p, lim, c);
            if (t == T_TAB) { ++p; return T_TAB; }
            if (t >= 0) return t;
        }
        ++p;
    }
}

// `[+-]?[0-9]{1,18}` over the whole field, else null (rcls.rs:132-134 non-strict cast).
PTX_HD int parse_int_field(const uint8_t* b, uint32_t& p, uint32_t lim, int64_t& out) {
    out = NULL_I64;
    if (p >= lim) return T_LIMIT;
    uint8_t c = b[p];
    bool neg = false;
    if (c == '-' || c == '+') {
        neg = (c == '-');
        ++p;
        if (p >= lim) return T_LIMIT;
        c = b[p];
    }
    uint64_t v = 0;
    uint32_t nd = 0;
    for (;;) {
        uint32_t d = (uint32_t)c - (uint32_t)'0';
        if (d > 9u) break;
        v = v * 10u + d;
        ++nd;
        ++p;
        if (p >= lim) return T_LIMIT;
        c = b[p];
    }
    int t = term_at(b, p, lim, c);
    if (t == T_LIMIT) return T_LIMIT;
    if (t >= 0) {
        if (nd >= 1 && nd <= 18) out = neg ? -(int64_t)v : (int64_t)v;
        if (t == T_TAB) ++p;
        return t;
    }
    return skip_field(b, p, lim);  // junk in an integer column -> null
}

// Parses columns 1..12 of the line starting at b[p].  Returns false if `lim` was hit
// before column 12 (or the end of line) was reached; the caller retries on the
// global-memory copy of the line.
PTX_HD bool parse_record(const uint8_t* b, uint32_t p, uint32_t lim, RecParse& r) {
    r.qlen = r.c7 = r.c8 = r.c9 = r.mapq = NULL_I64;
    r.path_pos = r.path_end = p;
    r.W = 0;
    r.vmin = INT64_MAX;
    r.vmax = -1;
    r.path_null = true;
    r.monotone = true;
    int t;
    {  // column 1: read id -> 96-bit hash
        IdHasher H;
        for (;;) {
            if (p >= lim) return false;
            uint8_t c = b[p];
            if (c <= '\r') {
                t = term_at(b, p, lim, c);
                if (t == T_LIMIT) return false;
                if (t >= 0) break;
            }
            H.byte(c);
            ++p;
        }
        r.h = H.finish();
        if (t == T_EOL) return true;
        ++p;
    }
    t = parse_int_field(b, p, lim, r.qlen);  // column 2
    if (t == T_LIMIT) return false;
    if (t == T_EOL) return true;
    for (int k = 0; k < 3; ++k) {  // columns 3,4,5
        t = skip_field(b, p, lim);
        if (t == T_LIMIT) return false;
        if (t == T_EOL) return true;
    }
    {  // column 6: walk
        r.path_pos = p;
        uint64_t v = 0;
        uint32_t nd = 0;
        int64_t prev = 0;
        bool inc = true, dec = true;
        for (;;) {
            if (p >= lim) return false;
            uint8_t c = b[p];
            uint32_t d = (uint32_t)c - (uint32_t)'0';
            if (d <= 9u) {
                v = v * 10u + d;
                ++nd;
                ++p;
                continue;
            }
            if (nd) {
                if (nd <= 18) {
                    int64_t m = (int64_t)v;
                    if (r.W) {
                        if (m <= prev) inc = false;
                        if (m >= prev) dec = false;
                    }
                    prev = m;
                    if (m < r.vmin) r.vmin = m;
                    if (m > r.vmax) r.vmax = m;
                    ++r.W;
                }
                v = 0;
                nd = 0;
            }
            t = term_at(b, p, lim, c);
            if (t == T_LIMIT) return false;
            if (t >= 0) break;
            ++p;
        }
        r.path_end = p;
        r.monotone = inc || dec;
        r.path_null = (r.path_end - r.path_pos == 1u) && (b[r.path_pos] == '*');
        if (t == T_EOL) return true;
        ++p;
    }
    t = parse_int_field(b, p, lim, r.c7);
    if (t == T_LIMIT) return false;
    if (t == T_EOL) return true;
    t = parse_int_field(b, p, lim, r.c8);
    if (t == T_LIMIT) return false;
    if (t == T_EOL) return true;
    t = parse_int_field(b, p, lim, r.c9);
    if (t == T_LIMIT) return false;
    if (t == T_EOL) return true;
    for (int k = 0; k < 2; ++k) {  // columns 10, 11
        t = skip_field(b, p, lim);
        if (t == T_LIMIT) return false;
        if (t == T_EOL) return true;
    }
    t = parse_int_field(b, p, lim, r.mapq);  // column 12
    if (t == T_LIMIT) return false;
    return true;
}

// Iterates the digit runs (<= 18 digits) of b[p,end).
struct WalkIter {
    const uint8_t* b;
    uint32_t p, end;
    PTX_HD bool next(int64_t& m) {
        for (;;) {
            while (p < end && ((uint32_t)b[p] - (uint32_t)'0') > 9u) ++p;
            if (p >= end) return false;
            uint64_t v = 0;
            uint32_t nd = 0;
            while (p < end) {
                uint32_t d = (uint32_t)b[p] - (uint32_t)'0';
                if (d > 9u) break;
                v = v * 10u + d;
                ++nd;
                ++p;
            }
            if (nd <= 18) { m = (int64_t)v; return true; }
        }
    }
};

// profile.rs:787-919 for one coverage-eligible read of species `label`.
// Sink concept:
//   uint32_t len(uint32_t g);  void add_bases(uint32_t g, int64_t v);
//   void set_bits(uint32_t g, int64_t lo, int64_t hi, uint32_t ln);   0 <= lo < hi <= ln
//   void trio(uint32_t a, uint32_t b, uint32_t c, int64_t s);         global node indices, read order
//   void error_start_gt_len(uint32_t label);
template <class Sink>
PTX_HD void cover_record(const uint8_t* b, const RecParse& r, uint32_t label, int64_t range_start, int64_t node_base,
                         Sink& sink) {
    if (r.W == 0) return;  // profile.rs:794
    const int64_t ps = r.c8, pe = r.c9;
    int64_t target = pe - ps;  // :800
    WalkIter it{b, r.path_pos, r.path_end};
    int64_t m;
    if (r.W == 1) {  // :811
        it.next(m);
        const uint32_t g = (uint32_t)(node_base + (m - range_start));
        if (target < 0) return;  // :821-827
        sink.add_bases(g, target);  // :829
        const uint32_t ln = sink.len(g);
        if (ps >= 0 && ps < pe && pe <= (int64_t)ln) sink.set_bits(g, ps, pe, ln);  // :832-835
        return;
    }
    int64_t seen = 0;
    uint32_t ga = 0, gb = 0;
    int64_t rla = 0, rlb = 0;
    for (uint32_t i = 0; i < r.W; ++i) {
        it.next(m);
        const uint32_t g = (uint32_t)(node_base + (m - range_start));
        const int64_t ln = (int64_t)sink.len(g);
        int64_t aln, lo, hi;
        if (i == 0) {
            if (ps > ln) { sink.error_start_gt_len(label); return; }  // :854 (a panic in the reference)
            aln = ln - ps;
            lo = ps;
            hi = ln;  // min(ps + aln, ln) == ln
        } else if (i == r.W - 1) {
            if (target < seen) target = seen;  // :858
            aln = target - seen;
            lo = 0;
            hi = aln < ln ? aln : ln;  // :871
        } else {
            aln = ln;  // :861
            lo = 0;
            hi = ln;
        }
        if (lo >= 0 && hi > lo) sink.set_bits(g, lo, hi, (uint32_t)ln);  // negative start wraps `as usize` -> empty
        seen += aln;                                                      // :878
        bool first = true;
        int64_t rl = aln;
        if (!r.monotone) {  // exact first-occurrence test (:879) against earlier walk positions
            WalkIter jt{b, r.path_pos, r.path_end};
            int64_t mj;
            for (uint32_t j = 0; j < i; ++j) {
                jt.next(mj);
                if (mj == m) {
                    first = false;
                    rl = (j == 0) ? (ln - ps) : ln;  // what the first occurrence added (:880)
                    break;
                }
            }
        }
        if (first) sink.add_bases(g, aln);  // :881
        if (i >= 2) sink.trio(ga, gb, g, rla + rlb + rl);  // :890-906
        ga = gb; rla = rlb;
        gb = g;  rlb = rl;
    }
}

}  // namespace ptx
