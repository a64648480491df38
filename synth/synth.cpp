// Synthetic pangenome graphs + GAF streams of the shapes BASELINE.json names.
// TEST / BENCH INFRASTRUCTURE - not part of the product library.
//
// Deterministic: every random draw is splitmix64(seed, stream, counter), so the
// bytes produced do not depend on the number of generator threads.
//
// Graph model: a species is a chain of sites; a site has 1
// allele w.p. 0.80, 2 w.p. 0.15, 3 w.p. 0.05.  Backbone (1-allele) nodes are
// 1+geom(mean `backbone_mean`, default 64) bp capped at 1024 (reference constants.rs:3 chop size);
// variant alleles are 1 bp (SNP) w.p. 0.5, else 1+geom(mean 8).  A strain path
// picks one allele per site by hash(strain, site) and skips 1 % of variant
// sites (deletion).  Single-strain species are chopped into 1024 bp nodes
// (build_eq1.rs:38-119).  Global node ids are assigned species by species,
// which is what species_range.txt encodes (sort_range.rs:27-38).
//
// GAF dialects: vg-giraffe short reads (12 columns + AS:i, dv:f, cs:Z) and
// GraphAligner long reads (12 columns + NM:i AS:f dv:f id:f, so that column 16
// is the identity read by gaf_filter.rs:31).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return mix64(mix64(seed ^ mix64(stream)) + ctr * 0xd1342543de82ef95ULL);
}
inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

struct Species {
    int64_t start = 0;  // 1-based global id of the first node
    std::vector<int32_t> len;
    std::vector<std::vector<uint32_t>> paths;  // local 0-based ids
    std::vector<double> cum_w;                 // cumulative strain weights
};

struct Synth {
    uint64_t seed = 0;
    double backbone_mean = 64.0;
    std::vector<Species> sp;
    std::vector<double> cum_species;  // Zipf(1) cumulative
    int64_t total_nodes = 0;
};

int32_t geom_len(uint64_t r, double mean, int32_t cap) {
    double u = u01(r);
    if (u <= 0) u = 1e-18;
    double v = -std::log(u) * mean;
    int64_t l = 1 + (int64_t)v;
    return (int32_t)std::min<int64_t>(l, cap);
}

void build_species(Synth& S, int s, int64_t n_nodes, int n_haps) {
    Species& P = S.sp[s];
    const uint64_t sd = S.seed;
    const uint64_t st = 0x1000 + (uint64_t)s;
    P.len.reserve(n_nodes);
    P.paths.assign(n_haps, {});
    if (n_haps == 1) {
        for (int64_t i = 0; i < n_nodes; ++i) {
            int32_t l = 1024;
            if (i == n_nodes - 1) l = 1 + (int32_t)(rnd(sd, st, 7) % 1024);
            P.len.push_back(l);
            P.paths[0].push_back((uint32_t)i);
        }
    } else {
        int64_t site = 0;
        while ((int64_t)P.len.size() < n_nodes) {
            uint64_t r = rnd(sd, st, (uint64_t)site * 8);
            double u = u01(r);
            int k = u < 0.80 ? 1 : (u < 0.95 ? 2 : 3);
            int64_t room = n_nodes - (int64_t)P.len.size();
            if (k > room) k = (int)room;
            uint32_t first = (uint32_t)P.len.size();
            for (int a = 0; a < k; ++a) {
                uint64_t ra = rnd(sd, st, (uint64_t)site * 8 + 1 + a);
                int32_t l;
                if (k == 1) l = geom_len(ra, S.backbone_mean, 1024);
                else l = (u01(ra) < 0.5) ? 1 : geom_len(mix64(ra), 8.0, 1024);
                P.len.push_back(l);
            }
            for (int h = 0; h < n_haps; ++h) {
                uint64_t rh = rnd(sd ^ 0xabcdef, st * 131 + h, (uint64_t)site);
                if (k > 1 && (rh % 100) == 0) continue;  // deletion
                // strains are related: allele 0 is the major allele
                int a = 0;
                if (k > 1) {
                    double ua = u01(mix64(rh));
                    a = ua < 0.7 ? 0 : (int)(1 + (mix64(rh ^ 0x55) % (k - 1)));
                }
                P.paths[h].push_back(first + a);
            }
            ++site;
        }
    }
    P.cum_w.resize(n_haps);
    double acc = 0;
    for (int h = 0; h < n_haps; ++h) {
        double w = -std::log(std::max(1e-12, u01(rnd(sd, st, 0xffff0000ULL + h))));  // Dirichlet(1)
        if (n_haps > 2 && (rnd(sd, st, 0xfffe0000ULL + h) % 4) != 0) w = 0;  // most strains absent
        if (h == 0 && n_haps > 0) w = std::max(w, 0.5);
        acc += w;
        P.cum_w[h] = acc;
    }
}

struct OutBuf {
    std::string s;
    inline void ch(char c) { s.push_back(c); }
    inline void str(const char* p) { s.append(p); }
    inline void num(int64_t v) {
        char tmp[24];
        int n = 0;
        bool neg = v < 0;
        uint64_t u = neg ? (uint64_t)(-v) : (uint64_t)v;
        do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (neg) s.push_back('-');
        while (n) s.push_back(tmp[--n]);
    }
    inline void frac6(double v) {  // 0.xxxxxx
        if (v >= 1.0) { str("1"); return; }
        if (v < 0) v = 0;
        int64_t q = (int64_t)(v * 1e6 + 0.5);
        if (q >= 1000000) { str("1"); return; }
        char tmp[40];
        snprintf(tmp, sizeof tmp, "0.%06lld", (long long)q);
        str(tmp);
    }
};

struct GafParams {
    int long_reads;       // 0 short (vg giraffe dialect), 1 long (GraphAligner dialect)
    int read_len;         // short read length (150)
    double long_mean;     // 15000
    double long_sigma;    // 0.3
    double p_unmapped;    // '*' path
    double p_star_c9;     // '*' in column 9
    double p_neg_single;  // single-node record with pe < ps
    double p_chimera;     // walk spanning two species -> "U"
    double p_dup_same;    // extra record re-using an id, same species
    double p_dup_other;   // extra record re-using an id, different species
    double p_secondary;   // long reads: second, worse alignment line
    double p_comment;     // '@' comment line before the record
    int id_pair_suffix;   // 1: ids S<sp>R<i>/1|2
};

int pick_species(const Synth& S, double u) {
    auto it = std::lower_bound(S.cum_species.begin(), S.cum_species.end(), u * S.cum_species.back());
    size_t i = it - S.cum_species.begin();
    return (int)std::min(i, S.sp.size() - 1);
}
int pick_strain(const Species& P, double u) {
    auto it = std::lower_bound(P.cum_w.begin(), P.cum_w.end(), u * P.cum_w.back());
    size_t i = it - P.cum_w.begin();
    return (int)std::min(i, P.cum_w.size() - 1);
}

// One alignment line.  `rid` is the record counter used for the id.
void emit_record(const Synth& S, const GafParams& gp, uint64_t gseed, uint64_t rec, int64_t rid, int sp_force,
                 bool worse, OutBuf& o) {
    auto R = [&](uint64_t k) { return rnd(gseed, 0x77 + (worse ? 1 : 0), rec * 32 + k); };
    int s = sp_force >= 0 ? sp_force : pick_species(S, u01(R(0)));
    const Species& P = S.sp[s];
    int h = pick_strain(P, u01(R(1)));
    const std::vector<uint32_t>& path = P.paths[h];
    int64_t rlen;
    if (gp.long_reads) {
        // lognormal with the requested mean: mu = ln(mean) - sigma^2/2
        double u1 = std::max(1e-12, u01(R(2))), u2 = u01(R(3));
        double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        rlen = (int64_t)std::exp(std::log(gp.long_mean) - 0.5 * gp.long_sigma * gp.long_sigma + gp.long_sigma * z);
        if (rlen < 1200) rlen = 1200;
    } else {
        rlen = gp.read_len;
    }
    // id
    o.ch('S'); o.num(s); o.ch('R'); o.num(rid >> (gp.id_pair_suffix ? 1 : 0));
    if (gp.id_pair_suffix) { o.ch('/'); o.ch((rid & 1) ? '2' : '1'); }
    o.ch('\t'); o.num(rlen); o.ch('\t');
    double ua = u01(R(4));
    if (ua < gp.p_unmapped) {  // vg: unmapped read
        o.str("0\t"); o.num(rlen); o.str("\t*\t*\t*\t*\t*\t*\t*\t0\n");
        return;
    }
    ua -= gp.p_unmapped;
    bool star9 = false, neg_single = false, chimera = false;
    if (ua < gp.p_star_c9) star9 = true;
    else if ((ua -= gp.p_star_c9) < gp.p_neg_single) neg_single = true;
    else if ((ua -= gp.p_neg_single) < gp.p_chimera) chimera = S.sp.size() > 1;

    size_t i0 = (size_t)(R(5) % path.size());
    int64_t l0 = P.len[path[i0]];
    int64_t ps = (int64_t)(R(6) % (uint64_t)l0);
    int64_t need = ps + rlen;
    int64_t plen = 0;
    size_t i1 = i0;
    while (i1 < path.size() && plen < need) { plen += P.len[path[i1]]; ++i1; }
    if (neg_single) { i1 = i0 + 1; plen = l0; }
    int64_t pe = std::min(need, plen);
    bool rev = (R(7) & 1) != 0;
    size_t w = i1 - i0;
    if (neg_single) {  // cf. the real line quoted at profile.rs:813 (start beyond end)
        if (ps == 0) ps = 1;
        pe = (int64_t)(R(13) % (uint64_t)ps);
    } else if (rev) {  // reversed walk: offsets are relative to the reversed walk
        int64_t tail = plen - pe;
        pe = plen - ps;
        ps = tail;
    }
    o.str("0\t"); o.num(rlen); o.str("\t+\t");
    for (size_t k = 0; k < w; ++k) {
        uint32_t loc = rev ? path[i1 - 1 - k] : path[i0 + k];
        o.ch(rev ? '<' : '>');
        o.num(P.start + loc);
    }
    if (chimera) {  // append two nodes of another species
        int s2 = (int)((s + 1 + R(8) % (S.sp.size() - 1)) % S.sp.size());
        const Species& Q = S.sp[s2];
        const std::vector<uint32_t>& qp = Q.paths[0];
        size_t j0 = (size_t)(R(9) % qp.size());
        for (size_t k = 0; k < 2 && j0 + k < qp.size(); ++k) { o.ch('>'); o.num(Q.start + qp[j0 + k]); plen += Q.len[qp[j0 + k]]; }
    }
    o.ch('\t'); o.num(plen); o.ch('\t'); o.num(ps); o.ch('\t');
    if (star9) o.ch('*'); else o.num(pe);
    int64_t span = neg_single ? 0 : pe - ps;
    double ident = worse ? 0.80 + 0.1 * u01(R(10)) : 0.97 + 0.03 * u01(R(10));
    int64_t matches = (int64_t)((double)span * ident);
    o.ch('\t'); o.num(matches); o.ch('\t'); o.num(span); o.ch('\t');
    int mapq = (u01(R(11)) < 0.8) ? 60 : (int)(R(12) % 60);
    if (worse) mapq = (int)(R(12) % 30);
    o.num(mapq);
    if (gp.long_reads) {
        o.str("\tNM:i:"); o.num(span - matches);
        o.str("\tAS:f:"); o.num(matches * 2 - span);
        o.str("\tdv:f:"); o.frac6(1.0 - ident);
        o.str("\tid:f:"); o.frac6(ident);
    } else {
        o.str("\tAS:i:"); o.num(matches + 10);
        o.str("\tdv:f:"); o.frac6(1.0 - ident);
        o.str("\tcs:Z::"); o.num(span);
    }
    o.ch('\n');
}

void gen_block(const Synth& S, const GafParams& gp, uint64_t gseed, int64_t r0, int64_t r1, OutBuf& o) {
    o.s.reserve((size_t)(r1 - r0) * (gp.long_reads ? 700 : 150));
    for (int64_t r = r0; r < r1; ++r) {
        uint64_t u = rnd(gseed, 0x99, (uint64_t)r);
        if (u01(u) < gp.p_comment) o.str("@CO\tsynthetic comment line\n");
        emit_record(S, gp, gseed, (uint64_t)r, r, -1, false, o);
        double ud = u01(mix64(u));
        if (ud < gp.p_dup_same + gp.p_dup_other) {
            // a second record re-using the id just emitted
            // (same species: re-emit with a different rng stream; other species: force one)
            int s_first = pick_species(S, u01(rnd(gseed, 0x77, (uint64_t)r * 32)));
            int s2 = s_first;
            if (ud >= gp.p_dup_same && S.sp.size() > 1) s2 = (s_first + 1) % (int)S.sp.size();
            size_t mark = o.s.size();
            emit_record(S, gp, gseed, (uint64_t)r, r, s2, true, o);
            // ids embed the species; rewrite the id so that it is byte-identical to the first record's
            std::string line = o.s.substr(mark);
            o.s.resize(mark);
            size_t tab = line.find('\t');
            OutBuf idb;
            idb.ch('S'); idb.num(s_first); idb.ch('R'); idb.num(r >> (gp.id_pair_suffix ? 1 : 0));
            if (gp.id_pair_suffix) { idb.ch('/'); idb.ch((r & 1) ? '2' : '1'); }
            o.s.append(idb.s);
            o.s.append(line, tab, std::string::npos);
        } else if (gp.long_reads && u01(mix64(u ^ 0x1234)) < gp.p_secondary) {
            int s_first = pick_species(S, u01(rnd(gseed, 0x77, (uint64_t)r * 32)));
            emit_record(S, gp, gseed, (uint64_t)r, r, s_first, true, o);
        }
    }
}

}  // namespace

extern "C" {

void* synth_create(uint64_t seed, int n_species, const int64_t* nodes_per_species, const int* haps_per_species,
                   double backbone_mean, int n_threads) {
    Synth* S = new Synth;
    S->seed = seed;
    S->backbone_mean = backbone_mean;
    S->sp.resize(n_species);
    int64_t off = 1;
    for (int s = 0; s < n_species; ++s) {
        S->sp[s].start = off;
        off += nodes_per_species[s];
    }
    S->total_nodes = off - 1;
    std::vector<std::thread> th;
    int nt = std::max(1, n_threads);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (int s = t; s < n_species; s += nt) build_species(*S, s, nodes_per_species[s], haps_per_species[s]);
        });
    for (auto& x : th) x.join();
    double acc = 0;
    S->cum_species.resize(n_species);
    for (int s = 0; s < n_species; ++s) { acc += 1.0 / (double)(s + 1); S->cum_species[s] = acc; }
    return S;
}
void synth_destroy(void* h) { delete (Synth*)h; }
int64_t synth_total_nodes(void* h) { return ((Synth*)h)->total_nodes; }
int64_t synth_species_start(void* h, int s) { return ((Synth*)h)->sp[s].start; }
int64_t synth_species_nodes(void* h, int s) { return (int64_t)((Synth*)h)->sp[s].len.size(); }
int synth_species_haps(void* h, int s) { return (int)((Synth*)h)->sp[s].paths.size(); }
const int32_t* synth_species_len(void* h, int s) { return ((Synth*)h)->sp[s].len.data(); }
int64_t synth_path_size(void* h, int s, int hap) { return (int64_t)((Synth*)h)->sp[s].paths[hap].size(); }
const uint32_t* synth_path_nodes(void* h, int s, int hap) { return ((Synth*)h)->sp[s].paths[hap].data(); }

// Generates records [r0, r1) of the stream `gseed`.  Returns a malloc'ed buffer
// (free with synth_free) and its size in *out_size.
char* synth_gaf(void* h, uint64_t gseed, int64_t r0, int64_t r1, const double* params, int n_threads,
                int64_t* out_size) {
    const Synth& S = *(Synth*)h;
    GafParams gp;
    gp.long_reads = (int)params[0];
    gp.read_len = (int)params[1];
    gp.long_mean = params[2];
    gp.long_sigma = params[3];
    gp.p_unmapped = params[4];
    gp.p_star_c9 = params[5];
    gp.p_neg_single = params[6];
    gp.p_chimera = params[7];
    gp.p_dup_same = params[8];
    gp.p_dup_other = params[9];
    gp.p_secondary = params[10];
    gp.p_comment = params[11];
    gp.id_pair_suffix = (int)params[12];
    const int64_t BLK = 1 << 15;
    int64_t nblk = (r1 - r0 + BLK - 1) / BLK;
    std::vector<OutBuf> bufs((size_t)nblk);
    int nt = std::max(1, n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (int64_t b = t; b < nblk; b += nt)
                gen_block(S, gp, gseed, r0 + b * BLK, std::min(r1, r0 + (b + 1) * BLK), bufs[(size_t)b]);
        });
    for (auto& x : th) x.join();
    size_t total = 0;
    for (auto& b : bufs) total += b.s.size();
    char* out = (char*)malloc(total + 1);
    size_t off = 0;
    for (auto& b : bufs) { memcpy(out + off, b.s.data(), b.s.size()); off += b.s.size(); std::string().swap(b.s); }
    out[total] = 0;
    *out_size = (int64_t)total;
    return out;
}
void synth_free(char* p) { free(p); }

}  // extern "C"
