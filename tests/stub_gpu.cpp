// TEST INFRASTRUCTURE ONLY - never built into or shipped with the product.
//
// A stand-in `libpantax_gpu.so` for machines WITHOUT a GPU: the part of the C ABI (include/pantax_gpu.h) that the host driver
// `pantax-gpu-profile` calls, answered by the C++ restatement of the reference (oracle/oracle_cpu.cpp, compiled in).  The CPU test
// suite points the real driver binary at it (LD_LIBRARY_PATH) so that the driver's own logic - file formats, streaming, TSV
// writers, the float tail, the strain-stage hand-off - is exercised end to end where no CUDA device exists.  The numbers it
// returns are the restatement's, so these tests say nothing about the kernels (the `-m gpu` tests do that against the real library);
// nothing under pantax_b200/ builds, loads or links this file.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../include/pantax_gpu.h"
#include "../pantax_b200/csrc/ptx_fxorder.h"

extern "C" {
void* orc_create(int n_threads);
void orc_destroy(void* h);
void orc_set_ranges(void* h, int S, const char* const* taxid, const int64_t* start, const int64_t* end);
int orc_set_graph(void* h, int s, const int64_t* nodes_len, int64_t n, const uint64_t* path_off, const uint64_t* path_nodes, int64_t H);
double orc_prepare_graphs(void* h);
int orc_run(void* h, const uint8_t* data, size_t n);
void orc_set_labels(void* h, const uint32_t* labels, int64_t n);
int64_t orc_n_records(void* h);
int orc_ids_unique(void* h);
void orc_labels(void* h, uint32_t* out);
void orc_record_fields(void* h, int64_t* read_len, int64_t* mapq);
void orc_species_counts(void* h, int64_t* out);
int orc_species_error(void* h, int s);
int64_t orc_trio_count(void* h, int s);
void orc_trio_table(void* h, int s, uint64_t* keys3, int64_t* len, uint32_t* owner);
void orc_node_bases(void* h, int s, int64_t* out);
void orc_node_cov(void* h, int s, uint64_t* out);
void orc_trio_bases(void* h, int s, int64_t* out);
void orc_path_sums(void* h, int s, int64_t* cov, int64_t* len);
void orc_hap_trio_counts(void* h, int s, int64_t* U, int64_t* nz);
int64_t orc_filter_gaf(const uint8_t* data, size_t n, uint64_t* out_line_off, int64_t cap);
}

struct StubGraph {
    bool present = false;
    std::vector<int64_t> len;
    std::vector<uint64_t> off{0}, nodes;
    std::vector<std::string> names;
};
struct ptx_ctx {
    std::vector<std::string> taxid;
    std::vector<int64_t> start, end;
    std::vector<StubGraph> g;
    std::vector<uint8_t> gaf;
    std::vector<uint32_t> labels;
    bool have_labels = false, committed = false;
    void* orc = nullptr;
    std::string err;
};

namespace {
int fail(ptx_ctx* c, int code, const std::string& m) { c->err = m; return code; }
// plain reading of a well-formed species GFA (profile.rs:466-545): enough for the test databases
int parse_gfa(const uint8_t* b, size_t n, StubGraph& g) {
    std::map<std::string, std::vector<uint64_t>> paths;
    size_t i = 0;
    int64_t idx = 0;
    while (i < n) {
        const void* nl = memchr(b + i, '\n', n - i);
        size_t e = nl ? (size_t)((const uint8_t*)nl - b) : n;
        std::string line((const char*)b + i, e - i);
        i = e + 1;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> p;
        size_t st = 0;
        for (size_t k = 0; k <= line.size(); ++k)
            if (k == line.size() || line[k] == '\t') { p.push_back(line.substr(st, k - st)); st = k + 1; }
        if (line[0] == 'S') {
            if (p.size() < 3) continue;
            if (atoll(p[1].c_str()) - 1 != idx) return PTX_E_NODE_ORDER;
            ++idx;
            if (p[2].empty()) return PTX_E_ZERO_LEN;
            g.len.push_back((int64_t)p[2].size());
        } else if (line[0] == 'W' || line[0] == 'P') {
            const bool w = p[0] == "W";
            std::string hap = p.size() > 1 ? p[1] : "";
            if (!w) hap = hap.substr(0, hap.find('#'));
            const std::string fld = w ? p.back() : (p.size() > 2 ? p[2] : "");
            auto& dst = paths[hap];
            for (size_t k = 0; k < fld.size();) {
                if (fld[k] >= '0' && fld[k] <= '9') {
                    uint64_t v = 0;
                    while (k < fld.size() && fld[k] >= '0' && fld[k] <= '9') v = v * 10 + (uint64_t)(fld[k++] - '0');
                    dst.push_back(v - 1);
                } else ++k;
            }
        }
    }
    for (auto& kv : paths) {
        g.names.push_back(kv.first);
        g.nodes.insert(g.nodes.end(), kv.second.begin(), kv.second.end());
        g.off.push_back((uint64_t)g.nodes.size());
    }
    g.present = true;
    return PTX_OK;
}
int run(ptx_ctx* c) {
    if (c->orc) orc_destroy(c->orc);
    c->orc = orc_create(2);
    std::vector<const char*> names;
    for (auto& t : c->taxid) names.push_back(t.c_str());
    orc_set_ranges(c->orc, (int)names.size(), names.data(), c->start.data(), c->end.data());
    if (c->committed)
        for (size_t s = 0; s < c->g.size(); ++s) {
            StubGraph& g = c->g[s];
            if (!g.present) continue;
            std::vector<uint64_t> nodes = g.nodes;
            if (nodes.empty()) nodes.push_back(0);
            if (orc_set_graph(c->orc, (int)s, g.len.data(), (int64_t)g.len.size(), g.off.data(), nodes.data(), (int64_t)g.off.size() - 1) != 0)
                return fail(c, PTX_E_NVERT_MISMATCH, "range size != number of nodes");
        }
    orc_prepare_graphs(c->orc);
    if (c->have_labels) orc_set_labels(c->orc, c->labels.data(), (int64_t)c->labels.size());
    static const uint8_t none = 0;
    orc_run(c->orc, c->gaf.empty() ? &none : c->gaf.data(), c->gaf.size());
    return PTX_OK;
}
int need(ptx_ctx* c, int s, bool graph) {
    if (!c || !c->orc) return PTX_E_STATE;
    if (s < 0 || s >= (int)c->g.size()) return PTX_E_RANGE;
    if (graph && !c->g[(size_t)s].present) return PTX_E_NO_GRAPH;
    return PTX_OK;
}
}  // namespace

extern "C" {
int ptx_create(int, ptx_ctx** out) { *out = new ptx_ctx; return PTX_OK; }
void ptx_destroy(ptx_ctx* c) { if (c) { if (c->orc) orc_destroy(c->orc); delete c; } }
const char* ptx_last_error(const ptx_ctx* c) { return c ? c->err.c_str() : ""; }
const char* ptx_version(void) { return "stub (tests/stub_gpu.cpp): the C++ restatement behind the C ABI, no GPU"; }
int ptx_set_ranges(ptx_ctx* c, int n, const char* const* taxid, const int64_t* st, const int64_t* en) {
    c->taxid.assign(taxid, taxid + n);
    c->start.assign(st, st + n);
    c->end.assign(en, en + n);
    c->g.assign((size_t)n, StubGraph());
    return PTX_OK;
}
int ptx_upload_graph(ptx_ctx* c, int s, const int64_t* len, int64_t n, const uint64_t* off, const uint64_t* nodes, int64_t H) {
    if (s < 0 || s >= (int)c->g.size()) return PTX_E_RANGE;
    if (c->end[(size_t)s] - c->start[(size_t)s] + 1 != n) return fail(c, PTX_E_NVERT_MISMATCH, "range size != number of nodes");
    StubGraph& g = c->g[(size_t)s];
    g = StubGraph();
    g.len.assign(len, len + n);
    g.off.assign(off, off + H + 1);
    g.nodes.assign(nodes, nodes + off[H]);
    g.names.assign((size_t)H, "");
    g.present = true;
    return PTX_OK;
}
int ptx_upload_graph_gfa(ptx_ctx* c, int s, const uint8_t* gfa, size_t n) {
    if (s < 0 || s >= (int)c->g.size()) return PTX_E_RANGE;
    StubGraph g;
    int rc = parse_gfa(gfa, n, g);
    if (rc) return fail(c, rc, "GFA");
    if (c->end[(size_t)s] - c->start[(size_t)s] + 1 != (int64_t)g.len.size()) return fail(c, PTX_E_NVERT_MISMATCH, "range size != number of nodes");
    c->g[(size_t)s] = g;
    return PTX_OK;
}
int ptx_species_graph(ptx_ctx* c, int s, int64_t* len, uint64_t* off, uint64_t* nodes) {
    if (s < 0 || s >= (int)c->g.size() || !c->g[(size_t)s].present) return PTX_E_NO_GRAPH;
    StubGraph& g = c->g[(size_t)s];
    if (len) memcpy(len, g.len.data(), g.len.size() * 8);
    if (off) memcpy(off, g.off.data(), g.off.size() * 8);
    if (nodes && !g.nodes.empty()) memcpy(nodes, g.nodes.data(), g.nodes.size() * 8);
    return PTX_OK;
}
int64_t ptx_species_path_steps(const ptx_ctx* c, int s) { return (int64_t)c->g[(size_t)s].nodes.size(); }
int ptx_species_path_name(ptx_ctx* c, int s, int64_t h, char* buf, size_t cap) {
    const std::string& nm = c->g[(size_t)s].names[(size_t)h];
    if (nm.size() + 1 > cap) return PTX_E_INVALID;
    memcpy(buf, nm.c_str(), nm.size() + 1);
    return (int)nm.size();
}
int ptx_commit_graphs(ptx_ctx* c) { c->committed = true; return PTX_OK; }
int ptx_host_alloc(size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? PTX_OK : PTX_E_NOMEM; }
int ptx_host_free(void* p) { free(p); return PTX_OK; }
int ptx_ingest_gaf(ptx_ctx* c, const uint8_t* bytes, size_t n, int) { c->gaf.insert(c->gaf.end(), bytes, bytes + n); return PTX_OK; }
int ptx_ingest_labels(ptx_ctx* c, const uint32_t* labels, int64_t n) { c->labels.assign(labels, labels + n); c->have_labels = true; return PTX_OK; }
int ptx_finalize(ptx_ctx* c) { return run(c); }
int64_t ptx_num_records(const ptx_ctx* c) { return c->orc ? orc_n_records(c->orc) : 0; }
int ptx_ids_unique(const ptx_ctx* c) { return c->orc ? orc_ids_unique(c->orc) : 1; }
int ptx_read_labels(ptx_ctx* c, uint32_t* labels) { if (!c->orc) return PTX_E_STATE; orc_labels(c->orc, labels); return PTX_OK; }
int ptx_species_counts(ptx_ctx* c, int64_t* counts) { if (!c->orc) return PTX_E_STATE; orc_species_counts(c->orc, counts); return PTX_OK; }
int ptx_equal_length(ptx_ctx* c, int* is_equal, int64_t* read_len) {
    if (!c->orc) return PTX_E_STATE;
    const int64_t R = orc_n_records(c->orc);
    std::vector<int64_t> rl((size_t)R + 1), mq((size_t)R + 1);
    std::vector<uint32_t> lab((size_t)R + 1);
    orc_record_fields(c->orc, rl.data(), mq.data());
    orc_labels(c->orc, lab.data());
    std::vector<int64_t> seen;  // profile.rs:311-322: distinct read_len (a null is a value) among the first 1000 non-U rows
    int64_t rows = 0;
    for (int64_t i = 0; i < R && rows < 1000; ++i) {
        if (lab[(size_t)i] == PTX_LABEL_UNCLASSIFIED) continue;
        ++rows;
        bool have = false;
        for (int64_t v : seen) have = have || v == rl[(size_t)i];
        if (!have) seen.push_back(rl[(size_t)i]);
    }
    *is_equal = seen.size() == 1;
    *read_len = seen.size() == 1 ? seen[0] : 0;
    return PTX_OK;
}
int64_t ptx_species_nodes(const ptx_ctx* c, int s) { return c->g[(size_t)s].present ? (int64_t)c->g[(size_t)s].len.size() : PTX_E_NO_GRAPH; }
int64_t ptx_species_paths(const ptx_ctx* c, int s) { return c->g[(size_t)s].present ? (int64_t)c->g[(size_t)s].off.size() - 1 : PTX_E_NO_GRAPH; }
int64_t ptx_species_trios(const ptx_ctx* c, int s) { return c->orc && c->g[(size_t)s].present ? orc_trio_count(c->orc, s) : PTX_E_NO_GRAPH; }
int ptx_node_bases(ptx_ctx* c, int s, int64_t* out) {
    int rc = need(c, s, true);
    if (rc) return rc;
    if (orc_species_error(c->orc, s)) return PTX_E_START_GT_LEN;
    orc_node_bases(c->orc, s, out);
    return PTX_OK;
}
int ptx_node_cov(ptx_ctx* c, int s, uint64_t* out) {
    int rc = need(c, s, true);
    if (rc) return rc;
    if (orc_species_error(c->orc, s)) return PTX_E_START_GT_LEN;
    orc_node_cov(c->orc, s, out);
    return PTX_OK;
}
int ptx_node_depth(ptx_ctx* c, int s, double* out) {
    int rc = need(c, s, true);
    if (rc) return rc;
    if (orc_species_error(c->orc, s)) return PTX_E_START_GT_LEN;
    const StubGraph& g = c->g[(size_t)s];
    std::vector<int64_t> b(g.len.size() + 1);
    orc_node_bases(c->orc, s, b.data());
    for (size_t i = 0; i < g.len.size(); ++i) out[i] = (double)b[i] / (double)g.len[i];
    return PTX_OK;
}
int ptx_trio_bases(ptx_ctx* c, int s, int64_t* out) {
    int rc = need(c, s, true);
    if (rc) return rc;
    orc_trio_bases(c->orc, s, out);
    return PTX_OK;
}
int ptx_trio_table(ptx_ctx* c, int s, uint64_t* keys3, int64_t* len, uint32_t* owner) {
    int rc = need(c, s, true);
    if (rc) return rc;
    const size_t T = (size_t)orc_trio_count(c->orc, s);
    std::vector<uint64_t> k(3 * T + 3);
    std::vector<int64_t> l(T + 1);
    std::vector<uint32_t> o(T + 1);
    orc_trio_table(c->orc, s, k.data(), l.data(), o.data());
    if (keys3) memcpy(keys3, k.data(), 3 * T * 8);
    if (len) memcpy(len, l.data(), T * 8);
    if (owner) memcpy(owner, o.data(), T * 4);
    return PTX_OK;
}
int ptx_trio_depth(ptx_ctx* c, int s, double* out) {
    int rc = need(c, s, true);
    if (rc) return rc;
    const size_t T = (size_t)orc_trio_count(c->orc, s);
    std::vector<uint64_t> k(3 * T + 3);
    std::vector<int64_t> l(T + 1), b(T + 1);
    std::vector<uint32_t> o(T + 1);
    orc_trio_table(c->orc, s, k.data(), l.data(), o.data());
    orc_trio_bases(c->orc, s, b.data());
    for (size_t t = 0; t < T; ++t) out[t] = (double)b[t] / (double)l[t];
    return PTX_OK;
}
int ptx_trio_ref_order(const uint64_t* path_off, const uint64_t* path_nodes, int64_t n_paths, const uint64_t* keys3, int64_t n_trios, uint64_t* order) {
    return ptx_fx::trio_ref_order(path_off, path_nodes, n_paths, keys3, n_trios, order) == 0 ? PTX_OK : PTX_E_INVALID;
}
int ptx_path_sums(ptx_ctx* c, int s, int64_t* cov, int64_t* len) {
    int rc = need(c, s, true);
    if (rc) return rc;
    const size_t H = c->g[(size_t)s].off.size() - 1;
    std::vector<int64_t> a(H + 1), b(H + 1);
    orc_path_sums(c->orc, s, a.data(), b.data());
    if (cov) memcpy(cov, a.data(), H * 8);
    if (len) memcpy(len, b.data(), H * 8);
    return PTX_OK;
}
int ptx_hap_trio_counts(ptx_ctx* c, int s, int64_t* U, int64_t* nz) {
    int rc = need(c, s, true);
    if (rc) return rc;
    const size_t H = c->g[(size_t)s].off.size() - 1;
    std::vector<int64_t> a(H + 1), b(H + 1);
    orc_hap_trio_counts(c->orc, s, a.data(), b.data());
    if (U) memcpy(U, a.data(), H * 8);
    if (nz) memcpy(nz, b.data(), H * 8);
    return PTX_OK;
}
int ptx_filter_gaf(ptx_ctx*, const uint8_t* bytes, size_t n, uint64_t* out_line_off, int64_t cap, int64_t* n_out) {
    *n_out = orc_filter_gaf(bytes, n, out_line_off, cap);
    return PTX_OK;
}
int ptx_stats_json(ptx_ctx*, char* buf, size_t cap) { snprintf(buf, cap, "{\"stub\": true}"); return PTX_OK; }
}  // extern "C"
