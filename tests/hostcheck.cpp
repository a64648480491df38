// CPU instantiation of pantax_b200/csrc/ptx_core.cuh (the per-record logic the CUDA
// kernels run) with a plain-array sink.  TEST HARNESS ONLY: lets `-m "not gpu"` tests
// check column parsing, classification, span arithmetic and the id hash against the
// oracle without a GPU.  It is not part of, nor linked into, libpantax_gpu.so.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../pantax_b200/csrc/ptx_core.cuh"
#include "../pantax_b200/csrc/ptx_fast.cuh"

using namespace ptx;

namespace {
struct HostSink {
    const uint32_t* len_;
    const uint64_t* bit_off;
    int64_t* bases;
    uint8_t* bytes;  // one byte per base
    const std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t>* tmap;
    int64_t* trio_bases;
    uint32_t* err;
    NodeInfo info(uint32_t g) const { return NodeInfo{len_[g], NI_TRIO_MID, bit_off[g]}; }
    void add_bases(uint32_t g, int64_t v) { bases[g] += v; }
    void set_bits(uint32_t, const NodeInfo& ni, int64_t lo, int64_t hi) {
        for (int64_t j = lo; j < hi; ++j) bytes[ni.bit_off + j] = 1;
    }
    void trio(uint32_t x, uint32_t y, uint32_t z, int64_t s, uint32_t /*y_flags*/) {
        uint32_t lo = x < z ? x : z, hi = x < z ? z : x;
        auto it = tmap->find(std::make_tuple(lo, y, hi));
        if (it != tmap->end()) trio_bases[it->second] += s;
    }
    void error_start_gt_len(uint32_t label) { err[label] |= 1u; }
};
struct HashKey {
    uint64_t lo; uint32_t hi;
    bool operator<(const HashKey& o) const { return lo != o.lo ? lo < o.lo : hi < o.hi; }
};
}  // namespace

extern "C" {

// Runs the whole per-record path on the CPU with the shared core.
// trio_keys: T x 3 canonical GLOBAL node idx.  Outputs are caller-allocated.
// `stage_lim`: if > 0, records are first parsed with that byte limit (emulating the
// shared-memory window) and re-parsed unbounded when parse_record reports overflow.
int hostcheck_run(const uint8_t* gaf, uint64_t n, int S, const int64_t* rstart, const int64_t* rend, const int64_t* node_base,
                  const uint32_t* order, int disjoint, int64_t N, const uint32_t* len, int64_t T, const uint32_t* trio_keys,
                  uint32_t stage_lim, int use_stash, uint32_t* labels, int64_t* n_records, int64_t* hist, int64_t* bases, uint64_t* cov,
                  int64_t* trio_bases, uint32_t* err, int* ids_unique, int64_t* n_overflow, int64_t* n_fast) {
    std::vector<uint32_t> sstart(S);
    for (int i = 0; i < S; ++i) sstart[i] = (uint32_t)rstart[order[i]];
    RangesView R{rstart, rend, node_base, order, sstart.data(), S, disjoint};
    std::vector<uint64_t> bit_off(N + 1, 0);
    for (int64_t i = 0; i < N; ++i) bit_off[i + 1] = bit_off[i] + len[i];
    std::vector<uint8_t> bytes(bit_off[N] + 1, 0);
    std::map<std::tuple<uint32_t, uint32_t, uint32_t>, uint32_t> tmap;
    for (int64_t t = 0; t < T; ++t) tmap[std::make_tuple(trio_keys[3 * t], trio_keys[3 * t + 1], trio_keys[3 * t + 2])] = (uint32_t)t;
    std::vector<uint8_t> buf(gaf, gaf + n);
    buf.insert(buf.end(), 64, '\n');  // padding, as the chunk buffers have
    struct Parsed { RecParse r; uint32_t label; uint32_t pos; bool eligible; uint32_t stash[8]; };
    std::vector<Parsed> recs;
    *n_overflow = 0;
    *n_fast = 0;
    // the word-wide fast path (ptx_fast.cuh) sees the text as the ingest kernel stages it: aligned words, a tab and
    // a newline bitmap built from 16-byte pieces, one sentinel word of all ones behind each bitmap
    const uint32_t f_lim = (uint32_t)(((n + 31) / 32) * 32 + 32);
    std::vector<uint32_t> f_words((f_lim + 128) / 4, 0x0a0a0a0au);
    memcpy(f_words.data(), gaf, n);
    std::vector<uint32_t> f_tab(f_lim / 32 + 2, 0xFFFFFFFFu), f_nl(f_lim / 32 + 2, 0xFFFFFFFFu);
    for (uint32_t pc = 0; pc < f_lim / 16; ++pc) {
        uint32_t nl16, tab16;
        classify16(f_words[4 * pc], f_words[4 * pc + 1], f_words[4 * pc + 2], f_words[4 * pc + 3], nl16, tab16);
        reinterpret_cast<uint16_t*>(f_nl.data())[pc] = (uint16_t)nl16;
        reinterpret_cast<uint16_t*>(f_tab.data())[pc] = (uint16_t)tab16;
    }
    for (uint64_t b = 0; b < f_lim; ++b) {  // the bitmaps are what a byte loop says
        const uint8_t c = reinterpret_cast<const uint8_t*>(f_words.data())[b];
        if ((((f_nl[b >> 5] >> (b & 31)) & 1u) != 0) != (c == '\n') || (((f_tab[b >> 5] >> (b & 31)) & 1u) != 0) != (c == '\t')) return 11;
    }
    uint64_t i = 0;
    while (i < n) {
        const uint8_t* nl = (const uint8_t*)memchr(buf.data() + i, '\n', buf.size() - i);
        uint64_t e = (uint64_t)(nl - buf.data());
        uint8_t c = buf[i];
        bool rec = c != '\n' && c != '@' && !(c == '\r' && buf[i + 1] == '\n');
        if (rec) {
            Parsed p;
            p.pos = (uint32_t)i;
            bool ok = false;
            if (stage_lim) {  // emulate the staged window: stage_lim bytes followed by the '\n' sentinel
                std::vector<uint8_t> win(buf.begin() + i, buf.begin() + std::min<uint64_t>(buf.size(), i + stage_lim));
                const uint32_t wl = (uint32_t)win.size();
                win.insert(win.end(), 16, '\n');
                ok = parse_record(win.data(), 0, wl, p.r, 1u, use_stash ? p.stash : nullptr, 1, 8);
                if (ok) {  // positions are relative to the window: rebase onto buf
                    p.r.path_pos += (uint32_t)i;
                    p.r.path_end += (uint32_t)i;
                } else {
                    ++*n_overflow;
                }
            }
            if (!ok) parse_record(buf.data(), (uint32_t)i, (uint32_t)buf.size() - 2, p.r, 1u, use_stash ? p.stash : nullptr, 1, 8);
            {   // the split parser used by the long-read kernel (parse_head / column 6 decoded separately / parse_tail)
                // must see the same columns as parse_record
                RecParse q;
                q.qlen = q.c7 = q.c8 = q.c9 = q.mapq = NULL_I64;
                uint32_t pp = (uint32_t)i;
                const int st = parse_head(buf.data(), pp, q, 1u);
                bool same = q.h.lo == p.r.h.lo && q.h.hi == p.r.h.hi && q.qlen == p.r.qlen;
                if (st == T_TAB) {
                    uint32_t e6 = pp;
                    while (buf[e6] != '\t' && buf[e6] != '\n') ++e6;  // what coop_walk_count finds
                    const int term = buf[e6] == '\t' ? T_TAB : T_EOL;
                    if (term == T_EOL && e6 > pp && buf[e6 - 1] == '\r') --e6;
                    same = same && pp == p.r.path_pos && e6 == p.r.path_end;
                    uint32_t pt = e6;
                    if (term == T_TAB) ++pt;
                    parse_tail(buf.data(), pt, term, q, 1u);
                }
                same = same && q.c7 == p.r.c7 && q.c8 == p.r.c8 && q.c9 == p.r.c9 && q.mapq == p.r.mapq;
                if (!same) return 9;
            }
            {   // fast path: whatever it accepts must be what parse_record read
                FastRec f;
                uint32_t fst[16];
                const Words fw{0u, f_words.data()}, ft{0u, f_tab.data()}, fn{0u, f_nl.data()};
                BitCursor nc;
                nc.seek(fn, (uint32_t)i);
                if (fast_parse(fw, ft, (uint32_t)i, nc.next(), f_lim, f, 1u, fst, 1, 16)) {
                    ++*n_fast;
                    auto same_int = [&](int64_t exact, uint32_t v, uint32_t bit) { return (f.nulls & bit) ? exact == NULL_I64 : exact == (int64_t)v; };
                    bool same = f.h.lo == p.r.h.lo && f.h.hi == p.r.h.hi && same_int(p.r.qlen, f.qlen, FN_QLEN) && same_int(p.r.c7, f.c7, FN_C7) &&
                                same_int(p.r.c8, f.c8, FN_C8) && same_int(p.r.c9, f.c9, FN_C9) && same_int(p.r.mapq, f.mapq, FN_MAPQ) &&
                                f.W == p.r.W && f.path_pos == p.r.path_pos && f.path_end == p.r.path_end && f.path_null == p.r.path_null;
                    if (same && f.W) same = (int64_t)f.vmin == p.r.vmin && (int64_t)f.vmax == p.r.vmax;
                    if (same) {
                        WalkIter it{buf.data(), p.r.path_pos, p.r.path_end};
                        int64_t m;
                        for (uint32_t k = 0; k < f.W && k < 16u && same; ++k) same = it.next(m) && m == (int64_t)fst[k];  // (only the first 16 are stashed)
                    }
                    if (!same) return 10;
                }
            }
            p.label = classify(R, p.r.W ? p.r.vmin : -1, p.r.W ? p.r.vmax : -1, R.sstart);
            // the two-level species search the ingest kernels use (pivots of the sorted starts) must give the same label, for every
            // pivot stride (1 = every start is a pivot ... S = a single pivot)
            for (int stride : {1, 2, 3, 7, S > 1 ? S / 2 : 1, S}) {
                if (stride < 1) continue;
                std::vector<uint32_t> pv;
                for (int i = 0; i < S; i += stride) pv.push_back(R.sstart[i]);
                if (classify_pivots(R, p.r.W ? p.r.vmin : -1, p.r.W ? p.r.vmax : -1, R.sstart, pv.data(), (int)pv.size(), stride) != p.label) return 12;
            }
            p.eligible = p.label != LABEL_U && !p.r.path_null && p.r.c7 != NULL_I64 && p.r.c8 != NULL_I64 && p.r.c9 != NULL_I64;
            recs.push_back(p);
        }
        i = e + 1;
    }
    *n_records = (int64_t)recs.size();
    std::map<HashKey, uint32_t> ds;
    bool dup = false, mixed = false;
    for (size_t k = 0; k < recs.size(); ++k) {
        const Parsed& p = recs[k];
        labels[k] = p.label;
        if (p.label == LABEL_U) continue;
        int64_t* h = hist + 4 * (int64_t)p.label;
        h[0] += 1;
        h[1] += p.r.qlen == NULL_I64 ? 0 : p.r.qlen;
        if (p.r.mapq != NULL_I64 && p.r.mapq >= 3 && p.r.mapq <= 60) { h[2] += 1; if (p.r.mapq == 60) h[3] += 1; }
        HashKey key{p.r.h.lo, p.r.h.hi};
        auto it = ds.find(key);
        if (it == ds.end()) ds[key] = p.eligible ? p.label : DS_NONE;
        else {
            dup = true;
            if (p.eligible) {
                if (it->second == DS_NONE) it->second = p.label;
                else if (it->second != p.label) { it->second = DS_MIXED; mixed = true; }
            }
        }
    }
    *ids_unique = dup ? 0 : 1;
    HostSink sink{len, bit_off.data(), bases, bytes.data(), &tmap, trio_bases, err};
    for (const Parsed& p : recs) {
        if (!p.eligible || node_base[p.label] < 0) continue;
        if (mixed && ds[HashKey{p.r.h.lo, p.r.h.hi}] == DS_MIXED) continue;
        RecParse rr = p.r;
        if (!use_stash) rr.monotone = false;  // k_ingest_s does not track it: cover_record must notice repeats on its own
        cover_record(buf.data(), rr, p.label, rstart[p.label], node_base[p.label], sink, 1u, p.stash, 1);
    }
    for (int64_t g = 0; g < N; ++g) {
        uint64_t c = 0;
        for (uint64_t j = bit_off[g]; j < bit_off[g + 1]; ++j) c += bytes[j];
        cov[g] = c;
    }
    return 0;
}

// id hash of one string (for collision / avalanche checks)
void hostcheck_idhash(const uint8_t* s, uint32_t n, uint64_t* lo, uint32_t* hi) {
    IdHasher H;
    for (uint32_t i = 0; i < n; ++i) H.byte(s[i]);
    IdHash h = H.finish();
    *lo = h.lo;
    *hi = h.hi;
}
uint32_t hostcheck_trio_hash(uint32_t a, uint32_t b, uint32_t c) { return trio_hash(a, b, c); }

}  // extern "C"
