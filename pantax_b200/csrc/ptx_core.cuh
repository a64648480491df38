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
        return finish_words();
    }
    // for callers that fed whole 4-byte words through mix() themselves (the zero-padded last one included) and set nbytes
    PTX_HD IdHash finish_words() {
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

// One 16-byte gather per walk node: everything the coverage pass needs to know about a node.
struct NodeInfo {
    uint32_t len;      // node length in bases
    uint32_t flags;    // NI_FULL: fully covered by some read; NI_TRIO_MID: middle node of at least one unique trio
    uint64_t bit_off;  // first bit of the node in the packed covered-base bitmap
};
constexpr uint32_t NI_FULL = 1u, NI_TRIO_MID = 2u;
// bits 8..31 of the flags: a 24-bit signature of the (lo, hi) neighbour pairs of the unique trios this node is the middle of.  A read
// window whose pair's bit is not set cannot be a unique trio: the probe of the trio table (a dependent gather, 13 % of k_apply's
// stall samples when every NI_TRIO_MID node probed) is skipped - 15 % of the nodes carry NI_TRIO_MID, 0.4 % of the windows hit.
PTX_HD uint32_t trio_sig_bit(uint32_t lo, uint32_t hi) {
    const uint32_t h = ((lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u)) * 0xC2B2AE35u;
    return 1u << (8u + (uint32_t)(((uint64_t)h * 24u) >> 32));
}

struct RangesView {
    const int64_t* start;      // [S] 1-based inclusive, file order (species_range.txt)
    const int64_t* end;        // [S]
    const int64_t* node_base;  // [S] offset of the species in the concatenated node arrays, -1 = no graph
    const uint32_t* order;     // [S] species indices sorted by start (used when disjoint)
    const uint32_t* sstart;    // [S] start[order[i]] as u32 (ids are < 2^32): the array the binary search walks
    int S;
    int disjoint;
};

// rcls.rs:253-257: FIRST range in file order with min>=start && max<=end.
// `sstart`: R.sstart or a copy of it in faster memory (the kernel keeps it in shared memory).
PTX_HD uint32_t classify(const RangesView& R, int64_t lo, int64_t hi, const uint32_t* sstart) {
    if (R.disjoint) {
        // ranges are pairwise disjoint: at most one can contain `lo`; binary search over the sorted starts
        if (lo < 0) return LABEL_U;  // no node in the walk (min = max = -1, rcls.rs:248): starts are >= 0
        if (R.S == 1) return (lo >= R.start[0] && hi <= R.end[0]) ? 0u : LABEL_U;
        const uint64_t ulo = (uint64_t)lo;
        int a = 0, b = R.S;  // last position with start <= lo
        while (a < b) {
            const int m = (a + b) >> 1;
            if ((uint64_t)sstart[m] <= ulo) a = m + 1; else b = m;
        }
        if (a == 0) return LABEL_U;
        const uint32_t s = R.order[a - 1];
        return (hi <= R.end[s]) ? s : LABEL_U;
    }
    for (int s = 0; s < R.S; ++s)
        if (lo >= R.start[s] && hi <= R.end[s]) return (uint32_t)s;
    return LABEL_U;
}

// The same with a two-level search: pv[i] = sstart[i * stride] (i < npv) sits in shared memory, so of the ~log2(S) dependent loads
// of the binary search only the last log2(stride) go to global memory (k_ingest_s with 1000 species: 14 % of its stall samples were
// this loop).  npv == 0: no pivots, plain classify.
constexpr int CLASSIFY_PIVOTS = 128;
PTX_HD uint32_t classify_pivots(const RangesView& R, int64_t lo, int64_t hi, const uint32_t* sstart, const uint32_t* pv, int npv, int stride) {
    if (!R.disjoint || R.S == 1 || npv == 0) return classify(R, lo, hi, sstart);
    if (lo < 0) return LABEL_U;
    const uint64_t ulo = (uint64_t)lo;
    int a = 0, b = npv;  // pivots <= lo
    while (a < b) {
        const int m = (a + b) >> 1;
        if ((uint64_t)pv[m] <= ulo) a = m + 1; else b = m;
    }
    if (a == 0) return LABEL_U;
    int a0 = (a - 1) * stride + 1, b0 = a * stride < R.S ? a * stride : R.S;  // sstart[(a-1)*stride] <= lo: the count of starts <= lo lies in [a0, b0]
    while (a0 < b0) {
        const int m = (a0 + b0) >> 1;
        if ((uint64_t)sstart[m] <= ulo) a0 = m + 1; else b0 = m;
    }
    const uint32_t s = R.order[a0 - 1];
    return (hi <= R.end[s]) ? s : LABEL_U;
}

struct RecParse {
    IdHash h;
    int64_t qlen, c7, c8, c9, mapq;  // NULL_I64 = null
    uint32_t path_pos, path_end;      // bytes [pos,end) of column 6
    uint32_t W;                       // walk nodes (digit runs of <= 18 digits)
    int64_t vmin, vmax;
    bool path_null;
    bool monotone;                    // strictly increasing or strictly decreasing ids => no repeats
    bool stashed;                     // all W node ids were saved in the caller's stash (W <= cap, ids < 2^32)
};

// Warp reconvergence points.  The per-column loops below have data-dependent trip counts; without an
// explicit __syncwarp after each column the lanes of a warp drift apart for the rest of the record
// (measured: 9.9 of 32 lanes active, profiles/r1a_ingest_ncu_summary.md).  `mask` = lanes that are
// parsing a record in this pass; every one of them executes every PTX_RECONVERGE / PTX_WARP_*.
#if defined(__CUDA_ARCH__)
#define PTX_RECONVERGE(m) __syncwarp(m)
#define PTX_WARP_MAX_U32(m, v) __reduce_max_sync((m), (v))
#define PTX_WARP_ALL(m, p) __all_sync((m), (p))
#else
#define PTX_RECONVERGE(m) ((void)(m))
#define PTX_WARP_MAX_U32(m, v) (v)
#define PTX_WARP_ALL(m, p) (p)
#endif

// All scanning loops below stop at a '\n' and never test a byte limit: the caller guarantees that
// b[lim] and b[lim+1] are '\n' (a sentinel behind the staged window / the newline padding of the chunk
// buffer).  A line end found at p >= lim-1 is therefore reported as T_LIMIT ("window exhausted") and
// the caller re-parses the record from global memory.

// terminator class of byte c at b[p] (c <= '\r'): T_TAB, T_EOL or -1
PTX_HD int term_at(const uint8_t* b, uint32_t p, uint8_t c) {
    if (c == '\t') return T_TAB;
    if (c == '\n') return T_EOL;
    if (c == '\r' && b[p + 1] == '\n') return T_EOL;
    return -1;
}

// Skip `n` whole fields in one loop (columns 3-5 and 10-11).  Returns T_TAB if n tabs were consumed.
PTX_HD int skip_fields(const uint8_t* b, uint32_t& p, int n) {
    for (;;) {
        const uint8_t c = b[p];
        if (c <= '\r') {  // '\t'=9 '\n'=10 '\r'=13
            if (c == '\t') {
                ++p;
                if (--n == 0) return T_TAB;
                continue;
            }
            if (c == '\n' || (c == '\r' && b[p + 1] == '\n')) return T_EOL;
        }
        ++p;
    }
}

// Advance to the end of the current field.  On T_TAB the tab is consumed.
PTX_HD int skip_field(const uint8_t* b, uint32_t& p) {
    int t;
    for (;;) {
        const uint8_t c = b[p];
        if (c <= '\r') {  // '\t'=9 '\n'=10 '\r'=13
            t = term_at(b, p, c);
            if (t >= 0) break;
        }
        ++p;
    }
    if (t == T_TAB) ++p;
    return t;
}

// `[+-]?[0-9]{1,18}` over the whole field, else null (rcls.rs:132-134 non-strict cast).
PTX_HD int parse_int_field(const uint8_t* b, uint32_t& p, int64_t& out) {
    out = NULL_I64;
    uint8_t c = b[p];
    bool neg = false;
    if (c == '-' || c == '+') {
        neg = (c == '-');
        c = b[++p];
    }
    uint32_t d = (uint32_t)c - (uint32_t)'0';
    uint32_t v32 = 0, nd = 0;
    while (d <= 9u && nd < 9u) {  // up to 9 digits fit 32-bit arithmetic
        v32 = v32 * 10u + d;
        ++nd;
        c = b[++p];
        d = (uint32_t)c - (uint32_t)'0';
    }
    uint64_t v = v32;
    while (d <= 9u) {  // long numbers: rare
        v = v * 10u + d;
        ++nd;
        c = b[++p];
        d = (uint32_t)c - (uint32_t)'0';
    }
    const int t = term_at(b, p, c);
    if (t >= 0) {
        if (nd >= 1 && nd <= 18) out = neg ? -(int64_t)v : (int64_t)v;
        if (t == T_TAB) ++p;
        return t;
    }
    return skip_field(b, p);  // junk in an integer column -> null
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

// Parses columns 1..12 of the line starting at b[p].  Returns false if the columns do not end inside
// the window b[0, lim) - the caller retries on the global-memory copy of the line.
// `stash` (may be null): node ids of the walk are saved at stash[i * stash_stride], i < stash_cap.
PTX_HD bool parse_record(const uint8_t* b, uint32_t p, uint32_t lim, RecParse& r, uint32_t mask, uint32_t* stash,
                         uint32_t stash_stride, uint32_t stash_cap) {
    r.qlen = r.c7 = r.c8 = r.c9 = r.mapq = NULL_I64;
    r.path_pos = r.path_end = p;
    r.W = 0;
    r.vmin = INT64_MAX;
    r.vmax = -1;
    r.path_null = true;
    r.monotone = true;
    r.stashed = false;
    int st;  // T_TAB: more columns follow; T_EOL: line ended
    // No per-column window test: every scanner stops at the sentinel newline behind the window, and after
    // a line end no further column is read - so p is compared with lim once, at the end.
#define PTX_CHECK_WINDOW() ((void)0)
    {  // column 1: read id -> 96-bit hash
        IdHasher H;
        for (;;) {
            const uint8_t c = b[p];
            if (c <= '\r') {
                st = term_at(b, p, c);
                if (st >= 0) break;
            }
            H.byte(c);
            ++p;
        }
        r.h = H.finish();
        if (st == T_TAB) ++p;
        PTX_CHECK_WINDOW();
    }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) { st = parse_int_field(b, p, r.qlen); PTX_CHECK_WINDOW(); }  // column 2
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = skip_fields(b, p, 3);  // columns 3,4,5
    PTX_RECONVERGE(mask);
    {  // column 6: walk.  The lanes of the warp advance one NODE per iteration, in lock-step.
        // Node ids of up to 9 digits (< 2^32; the reference's own species ranges are u32, profile.rs:547-551)
        // are tracked in 32-bit arithmetic; a longer run switches the column to the generic 64-bit scan below.
        const bool had6 = (st == T_TAB);
        bool done = !had6;
        bool inc = true, dec = true, fits = (stash != nullptr), big = false;
        uint32_t prev32 = 0, vmin32 = 0xFFFFFFFFu, vmax32 = 0;
        if (had6) r.path_pos = p;
        for (;;) {
            if (!done) {
                uint8_t c;
                uint32_t d;
                for (;;) {  // separator bytes ('>' '<' ...) up to the next digit or the end of the column
                    c = b[p];
                    d = (uint32_t)c - (uint32_t)'0';
                    if (d <= 9u) break;
                    if (c <= '\r') {
                        st = term_at(b, p, c);
                        if (st >= 0) { done = true; break; }
                    }
                    ++p;
                }
                if (!done) {
                    uint32_t v32 = 0, nd = 0;
                    while (d <= 9u && nd < 9u) {
                        v32 = v32 * 10u + d;
                        ++nd;
                        c = b[++p];
                        d = (uint32_t)c - (uint32_t)'0';
                    }
                    if (d <= 9u) {  // 10+ digits: rare
                        big = true;
                        while (d <= 9u) { c = b[++p]; d = (uint32_t)c - (uint32_t)'0'; }
                    }
                    if (r.W) {
                        if (v32 <= prev32) inc = false;
                        if (v32 >= prev32) dec = false;
                    }
                    prev32 = v32;
                    vmin32 = v32 < vmin32 ? v32 : vmin32;
                    vmax32 = v32 > vmax32 ? v32 : vmax32;
                    if (fits && r.W < stash_cap) stash[r.W * stash_stride] = v32;
                    else fits = false;
                    ++r.W;
                }
            }
            if (PTX_WARP_ALL(mask, done)) break;
        }
        if (had6) {
            r.path_end = p;
            if (big) {  // generic scan: ids up to 18 digits, longer runs dropped (rcls.rs:244 parse().ok())
                WalkIter it{b, r.path_pos, r.path_end};
                int64_t m, prev = 0;
                r.W = 0;
                inc = dec = true;
                fits = (stash != nullptr);
                while (it.next(m)) {
                    if (r.W) {
                        if (m <= prev) inc = false;
                        if (m >= prev) dec = false;
                    }
                    prev = m;
                    if (m < r.vmin) r.vmin = m;
                    if (m > r.vmax) r.vmax = m;
                    if (fits && r.W < stash_cap && m <= 0xFFFFFFFFll) stash[r.W * stash_stride] = (uint32_t)m;
                    else fits = false;
                    ++r.W;
                }
            } else if (r.W) {
                r.vmin = (int64_t)vmin32;
                r.vmax = (int64_t)vmax32;
            }
            r.monotone = inc || dec;
            r.stashed = fits;
            r.path_null = (r.path_end - r.path_pos == 1u) && (b[r.path_pos] == '*');
            if (st == T_TAB) ++p;
            PTX_CHECK_WINDOW();
        }
    }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) { st = parse_int_field(b, p, r.c7); PTX_CHECK_WINDOW(); }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) { st = parse_int_field(b, p, r.c8); PTX_CHECK_WINDOW(); }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) { st = parse_int_field(b, p, r.c9); PTX_CHECK_WINDOW(); }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = skip_fields(b, p, 2);  // columns 10, 11
    PTX_RECONVERGE(mask);
    if (st == T_TAB) {
        st = parse_int_field(b, p, r.mapq);  // column 12
        PTX_CHECK_WINDOW();
        if (st == T_TAB) st = T_EOL;  // anything after column 12 is ignored
    }
    PTX_RECONVERGE(mask);
#undef PTX_CHECK_WINDOW
    return p + 1u < lim;
}

// parse_record in three pieces, for callers that decode column 6 themselves (the warp-cooperative walk decode of
// the long-read kernel).  Same scanners, same results: parse_head = columns 1-5 (p ends at the first byte of
// column 6 if T_TAB is returned), parse_tail = columns 7-12 starting at p with the terminator `st` of column 6.
PTX_HD int parse_head(const uint8_t* b, uint32_t& p, RecParse& r, uint32_t mask) {
    int st;
    {
        IdHasher H;
        for (;;) {
            const uint8_t c = b[p];
            if (c <= '\r') {
                st = term_at(b, p, c);
                if (st >= 0) break;
            }
            H.byte(c);
            ++p;
        }
        r.h = H.finish();
        if (st == T_TAB) ++p;
    }
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = parse_int_field(b, p, r.qlen);  // column 2
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = skip_fields(b, p, 3);  // columns 3,4,5
    PTX_RECONVERGE(mask);
    return st;
}
PTX_HD void parse_tail(const uint8_t* b, uint32_t& p, int st, RecParse& r, uint32_t mask) {
    if (st == T_TAB) st = parse_int_field(b, p, r.c7);
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = parse_int_field(b, p, r.c8);
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = parse_int_field(b, p, r.c9);
    PTX_RECONVERGE(mask);
    if (st == T_TAB) st = skip_fields(b, p, 2);  // columns 10, 11
    PTX_RECONVERGE(mask);
    if (st == T_TAB) parse_int_field(b, p, r.mapq);  // column 12; anything after it is ignored
    PTX_RECONVERGE(mask);
}

// profile.rs:787-919 for one coverage-eligible read of species `label`.
// Sink concept:
//   NodeInfo info(uint32_t g);  void add_bases(uint32_t g, int64_t v);
//   void set_bits(uint32_t g, const NodeInfo& ni, int64_t lo, int64_t hi);   0 <= lo < hi <= ni.len
//   void trio(uint32_t a, uint32_t b, uint32_t c, int64_t s, uint32_t b_flags);   global node indices in read order; only called
//                                                               when b carries NI_TRIO_MID (b_flags: its flags word, for trio_sig_bit)
//   void error_start_gt_len(uint32_t label);
// `mask`: lanes of the warp that call this together.  The walk is processed in three phases so that the lanes
// execute the same code at the same time: the first node of every read, then the middle nodes in lock-step,
// then the last node of every read (walk order is preserved, which the non-stashed re-parse needs).
template <class Sink>
PTX_HD void cover_record(const uint8_t* b, const RecParse& r, uint32_t label, int64_t range_start, int64_t node_base,
                         Sink& sink, uint32_t mask, const uint32_t* stash, uint32_t stash_stride) {
    const int64_t ps = r.c8, pe = r.c9;
    int64_t target = pe - ps;  // :800
    WalkIter it{b, r.path_pos, r.path_end};
    const bool stashed = r.stashed;
    const uint32_t W = r.W;  // W == 0: profile.rs:794
    // rl of node i as the trio loop sees it (:897-900): what its FIRST occurrence in this read added.
    // Strictly monotone ids cannot repeat: either the parser said so (r.monotone), or it is noticed here while the walk
    // goes by (trend: 0 = one node seen, 1 = rising so far, 2 = falling so far, 3 = neither -> search the earlier nodes).
    uint32_t trend = 0;
    int64_t prev_m = 0;
    auto first_occurrence = [&](uint32_t i, int64_t m, int64_t ln, int64_t aln, int64_t& rl) -> bool {
        rl = aln;
        if (r.monotone) return true;
        const uint32_t step = m > prev_m ? 1u : (m < prev_m ? 2u : 3u);
        trend = trend == 0u ? step : (trend == step ? trend : 3u);
        prev_m = m;
        if (trend != 3u) return true;
        WalkIter jt{b, r.path_pos, r.path_end};
        for (uint32_t j = 0; j < i; ++j) {
            int64_t mj;
            if (stashed) mj = (int64_t)stash[j * stash_stride];
            else jt.next(mj);
            if (mj == m) {
                rl = (j == 0) ? (ln - ps) : ln;  // :880
                return false;
            }
        }
        return true;
    };
    // ---- first node
    bool ok = W >= 1;
    uint32_t gb = 0, ga = 0;
    int64_t rlb = 0, rla = 0, seen = 0;
    uint32_t fl_b = 0;  // flags of gb: NI_TRIO_MID = it is the middle node of some unique trio
    if (ok) {
        int64_t m;
        if (stashed) m = (int64_t)stash[0];
        else it.next(m);
        const uint32_t g = (uint32_t)(node_base + (m - range_start));
        const NodeInfo ni = sink.info(g);
        const int64_t ln = (int64_t)ni.len;
        if (W == 1) {  // :811
            if (target >= 0) {              // :821-827 (target < 0: the read is skipped)
                sink.add_bases(g, target);  // :829
                if (ps >= 0 && ps < pe && pe <= ln) sink.set_bits(g, ni, ps, pe);  // :832-835
            }
        } else if (ps > ln) {  // :854 (a panic in the reference): flag the species, drop the read
            sink.error_start_gt_len(label);
            ok = false;
        } else {
            const int64_t aln = ln - ps;                      // :856
            if (ps >= 0 && ln > ps) sink.set_bits(g, ni, ps, ln);  // negative start wraps `as usize` -> empty
            seen = aln;
            sink.add_bases(g, aln);  // :881 (position 0 is always a first occurrence)
            prev_m = m;
            gb = g;
            rlb = aln;
            fl_b = ni.flags;
        }
    }
    PTX_RECONVERGE(mask);
    // ---- middle nodes 1 .. W-2: fully covered (:861)
    const uint32_t nmid = (ok && W >= 3) ? W - 2 : 0;
    const uint32_t nmax = PTX_WARP_MAX_U32(mask, nmid);
    for (uint32_t j = 0; j < nmax; ++j) {
        if (j < nmid) {
            const uint32_t i = j + 1;
            int64_t m;
            if (stashed) m = (int64_t)stash[i * stash_stride];
            else it.next(m);
            const uint32_t g = (uint32_t)(node_base + (m - range_start));
            const NodeInfo ni = sink.info(g);
            const int64_t ln = (int64_t)ni.len;
            sink.set_bits(g, ni, 0, ln);
            seen += ln;  // :878
            int64_t rl;
            if (first_occurrence(i, m, ln, ln, rl)) sink.add_bases(g, ln);     // :879-882
            if (i >= 2 && (fl_b & NI_TRIO_MID)) sink.trio(ga, gb, g, rla + rlb + rl, fl_b);  // :890-906
            ga = gb; rla = rlb;
            gb = g;  rlb = rl;
            fl_b = ni.flags;
        }
        PTX_RECONVERGE(mask);
    }
    // ---- last node
    if (ok && W >= 2) {
        const uint32_t i = W - 1;
        int64_t m;
        if (stashed) m = (int64_t)stash[i * stash_stride];
        else it.next(m);
        const uint32_t g = (uint32_t)(node_base + (m - range_start));
        const NodeInfo ni = sink.info(g);
        const int64_t ln = (int64_t)ni.len;
        if (target < seen) target = seen;  // :858
        const int64_t aln = target - seen;
        const int64_t hi = aln < ln ? aln : ln;  // :871
        if (hi > 0) sink.set_bits(g, ni, 0, hi);
        int64_t rl;
        if (first_occurrence(i, m, ln, aln, rl)) sink.add_bases(g, aln);
        if (W >= 3 && (fl_b & NI_TRIO_MID)) sink.trio(ga, gb, g, rla + rlb + rl, fl_b);
    }
    PTX_RECONVERGE(mask);
}

}  // namespace ptx
