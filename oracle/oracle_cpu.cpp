// CPU oracle #2 for PanTax's alignment-to-abundance hot path - TEST INFRASTRUCTURE ONLY.
//
// A multithreaded C++17 restatement of the reference's algorithm
// (LuoGroup2023/PanTax v2.1.0, pantax/src/{rcls,profile,gaf_filter}.rs), written
// independently of oracle/pantax_oracle.py (the naive Python restatement) so that
// the two can be cross-checked, and fast enough to be the timed CPU baseline
// ("kind":"port") in bench.py.  Nothing under pantax_b200/ links or loads it.
//
// PARITY UNPINNED: the reference has no golden vectors / asserting tests for this
// path and cannot be compiled here (no Rust toolchain, crates un-vendored); see
// the header of oracle/pantax_oracle.py and DESIGN.md.  Pins: tests/golden/kat_*.json
// (hand-derived from SURVEY.md section 8c) + agreement with the Python restatement.
//
// Differences from a literal port, all in the reference's favour as a baseline:
//   * plain arrays with relaxed atomics instead of DashMap (profile.rs:774-785);
//   * covered bases counted once at the end instead of an O(len) recount on every
//     touch (profile.rs:844, 874);
//   * the id uniqueness set (profile.rs:369-378, a serial loop) is sharded and filled
//     in parallel;
//   * trio / path incidence kept sparse instead of dense DMatrix (profile.rs:686,
//     2705) so the stress sizes fit in memory.
// Threads default to all hardware threads, which is what the reference does
// (rayon/polars pools are never configured).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

constexpr uint32_t LABEL_U = 0xFFFFFFFFu;
constexpr int64_t NULL_I64 = INT64_MIN;

struct TrioKey {
    uint64_t a, b, c;
    bool operator==(const TrioKey& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct TrioHash {
    size_t operator()(const TrioKey& k) const {
        uint64_t h = k.a * 0x9e3779b97f4a7c15ULL;
        h ^= (k.b + 0x7f4a7c15ULL + (h << 6) + (h >> 2));
        h *= 0xff51afd7ed558ccdULL;
        h ^= (k.c + (h << 6) + (h >> 2));
        return (size_t)(h ^ (h >> 29));
    }
};

struct SpeciesGraph {
    bool present = false;
    std::vector<int64_t> len;
    std::vector<uint64_t> path_off;    // H+1
    std::vector<uint64_t> path_nodes;  // local ids
    // trio_nodes_info outputs (profile.rs:658-740), order = (owner hap, window position)
    std::unordered_map<TrioKey, uint32_t, TrioHash> trio_map;
    std::vector<TrioKey> trio_key;
    std::vector<int64_t> trio_len;
    std::vector<uint32_t> trio_owner;
    // get_node_abundances outputs (profile.rs:743-1026)
    std::vector<std::atomic<int64_t>> bases;
    std::vector<std::atomic<int64_t>> trio_bases;
    std::vector<uint64_t> bit_off;  // prefix of len, one byte per base as in profile.rs:780
    std::vector<std::atomic<uint8_t>> bits;
    std::vector<uint64_t> cov;
    std::vector<int64_t> path_cov, path_len, hapU, hapNZ;
    std::atomic<int> err_start_gt_len{0};
};

struct Rec {  // one GAF row (rcls.rs:127-136)
    const char* id;
    uint32_t id_len;
    const char* path;
    uint32_t path_len;  // 0 + path==nullptr => null
    int64_t read_len, c7, c8, c9, mapq;
    uint32_t label;
};

struct Oracle {
    int n_threads = 1;
    std::vector<std::string> taxid;
    std::vector<int64_t> rstart, rend;
    std::vector<SpeciesGraph> g;
    std::vector<Rec> recs;
    std::vector<int64_t> counts;  // S x 4
    std::vector<uint32_t> label_override;  // strain-only resume (profile.rs:3367-3385): species column from a file
    bool label_out_of_range = false;
    bool ids_unique = true;
    int64_t n_mixed_dropped = 0;
    double t_parse = 0, t_group = 0, t_cov = 0, t_stats = 0;
    std::atomic<int64_t> st_reads{0}, st_nodes{0}, st_windows{0}, st_hits{0};  // workload statistics for the roofline
};

template <class F>
void parallel_for(int nt, size_t n, F f) {
    if (nt <= 1 || n < 2) {
        f(0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    size_t chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        size_t a = std::min(n, t * chunk), b = std::min(n, a + chunk);
        th.emplace_back([=] { f(t, a, b); });
    }
    for (auto& x : th) x.join();
}

// `[+-]?[0-9]{1,18}` over the whole field, else null (polars strict=false cast, rcls.rs:132-134)
int64_t parse_int_field(const char* p, size_t n) {
    if (n == 0) return NULL_I64;
    size_t i = 0;
    bool neg = false;
    if (p[0] == '+' || p[0] == '-') { neg = p[0] == '-'; i = 1; }
    if (n - i == 0 || n - i > 18) return NULL_I64;
    int64_t v = 0;
    for (; i < n; ++i) {
        if (p[i] < '0' || p[i] > '9') return NULL_I64;
        v = v * 10 + (p[i] - '0');
    }
    return neg ? -v : v;
}

// maximal digit runs (regex \d+, rcls.rs:455 / profile.rs:769); runs > 18 digits dropped
template <class F>
void for_digit_runs(const char* p, size_t n, F f) {
    size_t i = 0;
    while (i < n) {
        if (p[i] >= '0' && p[i] <= '9') {
            size_t j = i;
            int64_t v = 0;
            while (j < n && p[j] >= '0' && p[j] <= '9') { v = v * 10 + (p[j] - '0'); ++j; }
            if (j - i <= 18) f(v);
            i = j;
        } else {
            ++i;
        }
    }
}

void parse_line(const char* p, size_t n, Rec& r) {
    // split on '\t' (rcls.rs:122,125); '*' => null (rcls.rs:124)
    const char* f[12];
    size_t fl[12];
    int nf = 0;
    size_t s = 0;
    for (size_t i = 0; i <= n && nf < 12; ++i) {
        if (i == n || p[i] == '\t') {
            f[nf] = p + s;
            fl[nf] = i - s;
            ++nf;
            s = i + 1;
        }
    }
    auto isnull = [&](int k) { return k >= nf || (fl[k] == 1 && f[k][0] == '*'); };
    r.id = f[0];
    r.id_len = (uint32_t)fl[0];
    r.read_len = isnull(1) ? NULL_I64 : parse_int_field(f[1], fl[1]);
    if (isnull(5)) { r.path = nullptr; r.path_len = 0; } else { r.path = f[5]; r.path_len = (uint32_t)fl[5]; }
    r.c7 = isnull(6) ? NULL_I64 : parse_int_field(f[6], fl[6]);
    r.c8 = isnull(7) ? NULL_I64 : parse_int_field(f[7], fl[7]);
    r.c9 = isnull(8) ? NULL_I64 : parse_int_field(f[8], fl[8]);
    r.mapq = isnull(11) ? NULL_I64 : parse_int_field(f[11], fl[11]);
    r.label = LABEL_U;
}

// rcls.rs:237-258: first range in file order with min>=start && max<=end
uint32_t classify(const Oracle& O, const Rec& r) {
    int64_t lo = -1, hi = -1;
    bool any = false;
    if (r.path)
        for_digit_runs(r.path, r.path_len, [&](int64_t v) {
            if (!any) { lo = hi = v; any = true; } else { lo = std::min(lo, v); hi = std::max(hi, v); }
        });
    for (size_t s = 0; s < O.rstart.size(); ++s)
        if (lo >= O.rstart[s] && hi <= O.rend[s]) return (uint32_t)s;
    return LABEL_U;
}

// profile.rs:658-740
void trio_nodes_info(SpeciesGraph& G) {
    size_t H = G.path_off.size() - 1;
    std::unordered_map<TrioKey, uint32_t, TrioHash> cnt;
    size_t total = 0;
    for (size_t h = 0; h < H; ++h) {
        size_t n = G.path_off[h + 1] - G.path_off[h];
        if (n >= 3) total += n - 2;
    }
    cnt.reserve(total / 8 + 16);
    auto key_at = [&](size_t h, size_t i) {
        const uint64_t* p = &G.path_nodes[G.path_off[h] + i];
        TrioKey k{p[0], p[1], p[2]};
        if (k.a > k.c) std::swap(k.a, k.c);  // :672-678
        return k;
    };
    for (size_t h = 0; h < H; ++h) {
        size_t n = G.path_off[h + 1] - G.path_off[h];
        for (size_t i = 0; i + 2 < n; ++i) ++cnt[key_at(h, i)];  // :689-702 with multiplicity
    }
    G.trio_map.clear();
    G.trio_key.clear();
    G.trio_len.clear();
    G.trio_owner.clear();
    for (size_t h = 0; h < H; ++h) {
        size_t n = G.path_off[h + 1] - G.path_off[h];
        for (size_t i = 0; i + 2 < n; ++i) {
            TrioKey k = key_at(h, i);
            if (cnt[k] == 1) {  // :709
                G.trio_map.emplace(k, (uint32_t)G.trio_key.size());
                G.trio_key.push_back(k);
                G.trio_len.push_back(G.len[k.a] + G.len[k.b] + G.len[k.c]);  // :712
                G.trio_owner.push_back((uint32_t)h);
            }
        }
    }
}

// profile.rs:787-919 for one read
void cover_read(SpeciesGraph& G, int64_t range_start, const Rec& r, std::vector<int64_t>& nodes,
                std::vector<int64_t>& aln_of, std::vector<int64_t>& rl, int64_t* stats) {
    nodes.clear();
    for_digit_runs(r.path, r.path_len, [&](int64_t v) { nodes.push_back(v - range_start); });  // :788-792
    if (nodes.empty()) return;                                                                  // :794
    const size_t W = nodes.size();
    stats[0] += 1;
    stats[1] += (int64_t)W;
    int64_t target = r.c9 - r.c8;  // :800
    const int64_t ps = r.c8, pe = r.c9;
    auto setbits = [&](int64_t nd, int64_t lo, int64_t hi) {  // [lo,hi) clipped as `as usize` ranges do
        if (lo < 0) return;                                    // negative start wraps -> empty range
        uint64_t base = G.bit_off[nd];
        for (int64_t j = lo; j < hi; ++j) G.bits[base + j].store(1, std::memory_order_relaxed);
    };
    rl.assign(W, 0);  // rl[i] = read_nodes_len[nodes[i]] as seen by the trio loop
    if (W == 1) {      // :811
        int64_t nd = nodes[0];
        if (target < 0) return;  // :821-827
        G.bases[nd].fetch_add(target, std::memory_order_relaxed);  // :829
        if (ps < pe && pe <= G.len[nd]) setbits(nd, ps, pe);       // :832-835
        return;                                                    // < 3 nodes, no trios
    }
    aln_of.assign(W, 0);
    int64_t seen = 0;
    for (size_t i = 0; i < W; ++i) {
        int64_t nd = nodes[i], ln = G.len[nd], aln, sidx = 0;
        if (i == 0) {
            if (ps > ln) { G.err_start_gt_len.store(1); return; }  // :854 panic
            aln = ln - ps;
            sidx = ps;
        } else if (i == W - 1) {
            if (target < seen) target = seen;  // :858
            aln = target - seen;
        } else {
            aln = ln;
        }
        setbits(nd, sidx, std::min(sidx + aln, ln));  // :871
        seen += aln;                                  // :878
        size_t first = i;
        for (size_t j = 0; j < i; ++j)
            if (nodes[j] == nd) { first = j; break; }
        if (first == i) {  // :879-882
            aln_of[i] = aln;
            G.bases[nd].fetch_add(aln, std::memory_order_relaxed);
        }
        rl[i] = aln_of[first];
    }
    // rl of an earlier occurrence is final only after the loop (it is the first occurrence's value already)
    if (W < 3 || G.trio_map.empty()) return;  // :886
    for (size_t i = 0; i + 2 < W; ++i) {
        TrioKey k{(uint64_t)nodes[i], (uint64_t)nodes[i + 1], (uint64_t)nodes[i + 2]};
        int64_t s = rl[i] + rl[i + 1] + rl[i + 2];  // :897-900
        if (k.a > k.c) std::swap(k.a, k.c);         // :902-904 (keys are canonical)
        auto it = G.trio_map.find(k);
        stats[2] += 1;
        if (it != G.trio_map.end()) { stats[3] += 1; G.trio_bases[it->second].fetch_add(s, std::memory_order_relaxed); }
    }
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void* orc_create(int n_threads) {
    Oracle* O = new Oracle;
    O->n_threads = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    return O;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
int orc_threads(void* h) { return ((Oracle*)h)->n_threads; }

void orc_set_ranges(void* h, int S, const char* const* taxid, const int64_t* start, const int64_t* end) {
    Oracle& O = *(Oracle*)h;
    O.taxid.assign(taxid, taxid + S);
    O.rstart.assign(start, start + S);
    O.rend.assign(end, end + S);
    O.g = std::vector<SpeciesGraph>(S);
}

// paths in hap-name order (BTreeMap, types.rs:54); node ids local 0-based
int orc_set_graph(void* h, int s, const int64_t* nodes_len, int64_t n, const uint64_t* path_off,
                  const uint64_t* path_nodes, int64_t H) {
    Oracle& O = *(Oracle*)h;
    SpeciesGraph& G = O.g[s];
    if (O.rend[s] - O.rstart[s] + 1 != n) return -1;  // profile.rs:2938 nvert
    G.present = true;
    G.len.assign(nodes_len, nodes_len + n);
    G.path_off.assign(path_off, path_off + H + 1);
    G.path_nodes.assign(path_nodes, path_nodes + path_off[H]);
    G.bit_off.resize(n + 1);
    uint64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) { G.bit_off[i] = acc; acc += (uint64_t)nodes_len[i]; }
    G.bit_off[n] = acc;
    return 0;
}

// graph-only preprocessing (profile.rs:658-740), parallel over species (profile.rs:3297)
double orc_prepare_graphs(void* h) {
    Oracle& O = *(Oracle*)h;
    double t0 = now();
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < O.n_threads; ++t)
        th.emplace_back([&] {
            for (;;) {
                size_t s = next.fetch_add(1);
                if (s >= O.g.size()) break;
                if (O.g[s].present) trio_nodes_info(O.g[s]);
            }
        });
    for (auto& x : th) x.join();
    return now() - t0;
}

// The timed hot path: GAF bytes -> labels, species counts, node coverage, strain statistics.
int orc_run(void* h, const uint8_t* data, size_t n) {
    Oracle& O = *(Oracle*)h;
    const char* d = (const char*)data;
    const int nt = O.n_threads;
    const size_t S = O.rstart.size();
    double t0 = now();
    // ---- a1/a2: parse + classify, parallel over byte ranges snapped to newlines (rcls.rs:119-146, 306-323)
    std::vector<std::vector<Rec>> part(nt);
    parallel_for(nt, n, [&](int t, size_t a, size_t b) {
        if (a > 0) { while (a < b && d[a - 1] != '\n') ++a; }  // start at a line start
        std::vector<Rec>& out = part[t];
        size_t i = a;
        while (i < b) {
            const char* nl = (const char*)memchr(d + i, '\n', n - i);
            size_t e = nl ? (size_t)(nl - d) : n;
            size_t len = e - i;
            if (len && d[e - 1] == '\r') --len;
            if (len && d[i] != '@') {
                Rec r;
                parse_line(d + i, len, r);
                r.label = classify(O, r);
                out.push_back(r);
            }
            i = e + 1;
        }
    });
    O.recs.clear();
    size_t total = 0;
    for (auto& p : part) total += p.size();
    O.recs.reserve(total);
    for (auto& p : part) { O.recs.insert(O.recs.end(), p.begin(), p.end()); std::vector<Rec>().swap(p); }
    const size_t R = O.recs.size();
    if (!O.label_override.empty()) {  // profile.rs:3381-3385: hstack of the GAF columns and the species column
        if (O.label_override.size() < R) return -2;
        O.label_out_of_range = false;
        for (size_t i = 0; i < R; ++i) {
            Rec& r = O.recs[i];
            uint32_t lab = O.label_override[i];
            if (lab != LABEL_U && r.path) {  // a walk outside the labelled species' graph: local ids would index out of it
                int64_t lo = 0, hi = 0;
                bool any = false;
                for_digit_runs(r.path, r.path_len, [&](int64_t v) {
                    if (!any) { lo = hi = v; any = true; } else { lo = std::min(lo, v); hi = std::max(hi, v); }
                });
                if (any && (lo < O.rstart[lab] || hi > O.rend[lab])) { O.label_out_of_range = true; lab = LABEL_U; }
            }
            r.label = lab;
        }
    }
    // ---- a3: species counts (profile.rs:208-297, integer part)
    std::vector<std::vector<int64_t>> pc(nt, std::vector<int64_t>(S * 4, 0));
    parallel_for(nt, R, [&](int t, size_t a, size_t b) {
        int64_t* c = pc[t].data();
        for (size_t i = a; i < b; ++i) {
            const Rec& r = O.recs[i];
            if (r.label == LABEL_U) continue;
            int64_t* q = c + 4 * (size_t)r.label;
            q[0] += 1;
            q[1] += r.read_len == NULL_I64 ? 0 : r.read_len;
            if (r.mapq != NULL_I64 && r.mapq >= 3 && r.mapq <= 60) { q[2] += 1; if (r.mapq == 60) q[3] += 1; }
        }
    });
    O.counts.assign(S * 4, 0);
    for (auto& c : pc) for (size_t i = 0; i < S * 4; ++i) O.counts[i] += c[i];
    O.t_parse = now() - t0;
    t0 = now();
    // ---- a4: id uniqueness over all non-U rows (profile.rs:369-378); mixed-species groups (406-437)
    constexpr int NSH = 256;
    struct Shard { std::mutex m; std::unordered_map<std::string_view, uint32_t> first_species; };
    std::vector<Shard> shards(NSH);
    std::atomic<bool> dup{false};
    constexpr uint32_t NONE = 0xFFFFFFFEu, MIXED = 0xFFFFFFFDu;
    parallel_for(nt, R, [&](int, size_t a, size_t b) {
        std::hash<std::string_view> H;
        for (size_t i = a; i < b; ++i) {
            const Rec& r = O.recs[i];
            if (r.label == LABEL_U) continue;
            std::string_view id(r.id, r.id_len);
            bool eligible = r.path && r.c7 != NULL_I64 && r.c8 != NULL_I64 && r.c9 != NULL_I64;
            Shard& sh = shards[H(id) % NSH];
            std::lock_guard<std::mutex> lk(sh.m);
            auto it = sh.first_species.find(id);
            if (it == sh.first_species.end()) {
                sh.first_species.emplace(id, eligible ? r.label : NONE);
            } else {
                dup.store(true, std::memory_order_relaxed);
                if (eligible) {
                    if (it->second == NONE) it->second = r.label;
                    else if (it->second != r.label) it->second = MIXED;
                }
            }
        }
    });
    O.ids_unique = !dup.load();
    O.t_group = now() - t0;
    t0 = now();
    // ---- a7: node coverage (profile.rs:743-1026), parallel over reads
    for (auto& G : O.g) {
        if (!G.present) continue;
        size_t N = G.len.size();
        G.bases = std::vector<std::atomic<int64_t>>(N);
        G.trio_bases = std::vector<std::atomic<int64_t>>(G.trio_key.size());
        G.bits = std::vector<std::atomic<uint8_t>>(G.bit_off[N]);
        G.err_start_gt_len = 0;
    }
    std::atomic<int64_t> dropped{0};
    O.st_reads = 0; O.st_nodes = 0; O.st_windows = 0; O.st_hits = 0;
    parallel_for(nt, R, [&](int, size_t a, size_t b) {
        std::hash<std::string_view> H;
        std::vector<int64_t> nodes, aln_of, rl;
        int64_t stats[4] = {0, 0, 0, 0};
        for (size_t i = a; i < b; ++i) {
            const Rec& r = O.recs[i];
            if (r.label == LABEL_U) continue;
            if (!(r.path && r.c7 != NULL_I64 && r.c8 != NULL_I64 && r.c9 != NULL_I64)) continue;  // :380-399
            SpeciesGraph& G = O.g[r.label];
            if (!G.present) continue;
            if (!O.ids_unique) {
                std::string_view id(r.id, r.id_len);
                Shard& sh = shards[H(id) % NSH];
                if (sh.first_species.find(id)->second == MIXED) { dropped.fetch_add(1); continue; }  // :415-416
            }
            cover_read(G, O.rstart[r.label], r, nodes, aln_of, rl, stats);
        }
        O.st_reads += stats[0]; O.st_nodes += stats[1]; O.st_windows += stats[2]; O.st_hits += stats[3];
    });
    O.n_mixed_dropped = dropped.load();
    O.t_cov = now() - t0;
    t0 = now();
    // ---- covered bases per node (profile.rs:1018-1023), path sums (2705-2729), hap trio counts (1112-1135)
    for (auto& G : O.g) {
        if (!G.present) continue;
        size_t N = G.len.size();
        G.cov.assign(N, 0);
        parallel_for(nt, N, [&](int, size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                uint64_t c = 0;
                for (uint64_t j = G.bit_off[i]; j < G.bit_off[i + 1]; ++j) c += G.bits[j].load(std::memory_order_relaxed);
                G.cov[i] = c;
            }
        });
        size_t H = G.path_off.size() - 1;
        G.path_cov.assign(H, 0);
        G.path_len.assign(H, 0);
        G.hapU.assign(H, 0);
        G.hapNZ.assign(H, 0);
        parallel_for(std::min<int>(nt, (int)H), H, [&](int, size_t a, size_t b) {
            std::vector<uint8_t> mark(N);
            for (size_t p = a; p < b; ++p) {
                std::fill(mark.begin(), mark.end(), 0);
                int64_t sc = 0, sl = 0;
                for (uint64_t k = G.path_off[p]; k < G.path_off[p + 1]; ++k) {
                    uint64_t v = G.path_nodes[k];
                    if (!mark[v]) { mark[v] = 1; sc += (int64_t)G.cov[v]; sl += G.len[v]; }  // binary incidence :2706-2711
                }
                G.path_cov[p] = sc;
                G.path_len[p] = sl;
            }
        });
        for (size_t t = 0; t < G.trio_key.size(); ++t) {
            G.hapU[G.trio_owner[t]] += 1;
            if (G.trio_bases[t].load() > 0) G.hapNZ[G.trio_owner[t]] += 1;  // :1129-1135
        }
    }
    O.t_stats = now() - t0;
    return 0;
}

void orc_set_labels(void* h, const uint32_t* labels, int64_t n) { ((Oracle*)h)->label_override.assign(labels, labels + n); }
int orc_label_out_of_range(void* h) { return ((Oracle*)h)->label_out_of_range ? 1 : 0; }
int64_t orc_n_records(void* h) { return (int64_t)((Oracle*)h)->recs.size(); }
int orc_ids_unique(void* h) { return ((Oracle*)h)->ids_unique ? 1 : 0; }
int64_t orc_mixed_dropped(void* h) { return ((Oracle*)h)->n_mixed_dropped; }
void orc_times(void* h, double* t4) {
    Oracle& O = *(Oracle*)h;
    t4[0] = O.t_parse; t4[1] = O.t_group; t4[2] = O.t_cov; t4[3] = O.t_stats;
}
// [covered reads, their walk nodes, trio windows probed, probes that hit a unique trio]
void orc_workload_stats(void* h, int64_t* out4) {
    Oracle& O = *(Oracle*)h;
    out4[0] = O.st_reads; out4[1] = O.st_nodes; out4[2] = O.st_windows; out4[3] = O.st_hits;
}
void orc_labels(void* h, uint32_t* out) {
    Oracle& O = *(Oracle*)h;
    for (size_t i = 0; i < O.recs.size(); ++i) out[i] = O.recs[i].label;
}
void orc_record_fields(void* h, int64_t* read_len, int64_t* mapq) {
    Oracle& O = *(Oracle*)h;
    for (size_t i = 0; i < O.recs.size(); ++i) { read_len[i] = O.recs[i].read_len; mapq[i] = O.recs[i].mapq; }
}
void orc_species_counts(void* h, int64_t* out) {
    Oracle& O = *(Oracle*)h;
    memcpy(out, O.counts.data(), O.counts.size() * sizeof(int64_t));
}
int orc_species_error(void* h, int s) { return ((Oracle*)h)->g[s].err_start_gt_len.load(); }
int64_t orc_trio_count(void* h, int s) { return (int64_t)((Oracle*)h)->g[s].trio_key.size(); }
void orc_trio_table(void* h, int s, uint64_t* keys3, int64_t* len, uint32_t* owner) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    for (size_t t = 0; t < G.trio_key.size(); ++t) {
        keys3[3 * t] = G.trio_key[t].a; keys3[3 * t + 1] = G.trio_key[t].b; keys3[3 * t + 2] = G.trio_key[t].c;
        len[t] = G.trio_len[t];
        owner[t] = G.trio_owner[t];
    }
}
void orc_node_bases(void* h, int s, int64_t* out) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    for (size_t i = 0; i < G.bases.size(); ++i) out[i] = G.bases[i].load();
}
void orc_node_cov(void* h, int s, uint64_t* out) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    memcpy(out, G.cov.data(), G.cov.size() * sizeof(uint64_t));
}
void orc_trio_bases(void* h, int s, int64_t* out) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    for (size_t i = 0; i < G.trio_bases.size(); ++i) out[i] = G.trio_bases[i].load();
}
void orc_path_sums(void* h, int s, int64_t* cov, int64_t* len) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    memcpy(cov, G.path_cov.data(), G.path_cov.size() * sizeof(int64_t));
    memcpy(len, G.path_len.data(), G.path_len.size() * sizeof(int64_t));
}
void orc_hap_trio_counts(void* h, int s, int64_t* U, int64_t* nz) {
    SpeciesGraph& G = ((Oracle*)h)->g[s];
    memcpy(U, G.hapU.data(), G.hapU.size() * sizeof(int64_t));
    memcpy(nz, G.hapNZ.data(), G.hapNZ.size() * sizeof(int64_t));
}

// gaf_filter.rs:22-97.  Writes the byte offsets of the kept lines (file order, first
// qualifying line per read id) to out_line_off; returns how many.
int64_t orc_filter_gaf(const uint8_t* data, size_t n, uint64_t* out_line_off, int64_t cap) {
    const char* d = (const char*)data;
    struct PR { size_t off; std::string_view id; int32_t matches, mapq, span; double ident; };
    std::vector<PR> recs;
    size_t i = 0;
    auto parse_i32 = [](std::string_view f, int32_t& out) {
        if (f.empty()) return false;
        size_t k = 0; bool neg = false;
        if (f[0] == '+' || f[0] == '-') { neg = f[0] == '-'; k = 1; }
        if (k == f.size()) return false;
        int64_t v = 0;
        for (; k < f.size(); ++k) {
            if (f[k] < '0' || f[k] > '9') return false;
            v = v * 10 + (f[k] - '0');
            if (v > (int64_t)INT32_MAX + 1) return false;
        }
        v = neg ? -v : v;
        if (v < INT32_MIN || v > INT32_MAX) return false;
        out = (int32_t)v;
        return true;
    };
    while (i < n) {
        const char* nl = (const char*)memchr(d + i, '\n', n - i);
        size_t e = nl ? (size_t)(nl - d) : n;
        size_t a = i, b = e;
        auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; };
        while (a < b && ws(d[a])) ++a;      // gaf_filter.rs:23 line.trim()
        while (b > a && ws(d[b - 1])) --b;
        std::vector<std::string_view> f;
        size_t s = a;
        for (size_t k = a; k <= b; ++k)
            if (k == b || d[k] == '\t') { f.emplace_back(d + s, k - s); s = k + 1; }
        if (f.size() >= 16) {
            PR r;
            r.off = i;
            r.id = f[0];
            std::string_view idf = f[15];
            size_t c = idf.rfind(':');
            std::string tail(c == std::string_view::npos ? idf : idf.substr(c + 1));
            char* endp = nullptr;
            r.ident = strtod(tail.c_str(), &endp);
            bool okf = !tail.empty() && endp && *endp == 0 && tail.find_first_not_of("+-0123456789.eE") == std::string::npos;
            int32_t c3, c4;
            if (okf && parse_i32(f[9], r.matches) && parse_i32(f[11], r.mapq) && parse_i32(f[3], c4) && parse_i32(f[2], c3)) {
                r.span = (int32_t)((uint32_t)c4 - (uint32_t)c3);  // i32 subtraction wraps in the reference's release build
                recs.push_back(r);
            }
        }
        i = e + 1;
    }
    std::unordered_map<std::string_view, std::pair<int32_t, double>> best;
    for (auto& r : recs) {
        auto it = best.find(r.id);
        if (it == best.end()) best.emplace(r.id, std::make_pair(r.matches, r.ident));
        else if (r.matches > it->second.first || (r.matches == it->second.first && r.ident > it->second.second))
            it->second = {r.matches, r.ident};
    }
    std::unordered_set<std::string_view> written;
    int64_t k = 0;
    for (auto& r : recs) {
        if (!(r.mapq > 20 && r.span > 1000)) continue;
        auto& b = best[r.id];
        if (r.matches == b.first && r.ident == b.second && written.insert(r.id).second) {
            if (k < cap) out_line_off[k] = r.off;
            ++k;
        }
    }
    return k;
}

}  // extern "C"
