// pantax-gpu-profile: host driver of the PanTax profiling stage over the C ABI (include/pantax_gpu.h).
//
// The reference's host side is Rust (profile.rs::profile, profile.rs:3325-3436); no Rust toolchain exists in
// the build image, so the host glue is C++17.  It honours the reference's intermediate FILE FORMATS:
//   <db>/species_range.txt            taxid \t start \t end \t is_pan        (sort_range.rs:35-38, zip.rs:308)
//   <db>/species_genomes_stats.txt    taxid \t avg_len                      (stat.rs:131-139)
//   <db>/species_graph_info/<t>.bin   bincode 1.3 of types.rs:51-55 Graph   (zip.rs:185, read at zip.rs:236-247)
//   <db>/species_gfa/<t>.gfa          text GFA fallback                     (profile.rs:466-545, 2923-2927)
//   <gaf>                             gfa_mapped.gaf                        (utils.rs:52)
// and writes
//   <wd>/reads_classification.tsv     read_id \t mapq \t species \t read_len, no header   (profile.rs:3337-3351)
//   <wd>/species_abundance.txt        species_taxid predicted_abundance predicted_coverage  (profile.rs:338-347)
//   <wd>/strain_inputs/<t>.nodes.tsv / .paths.tsv   what optimize_otu hands to the ILP (profile.rs:2936-2967):
//        node depth, covered bases; per path: unique-trio fraction, frequencies_mean, path_cov_ratio, kept by
//        first_filter_paths; <wd>/strain_graphs/<t>.bin (plain bincode Graph) when the species' graph was not a plain .bin in the db.
//        `python -m pantax_b200.strain_tail --db DB --wd WD` reads these and carries the run to strain_abundance.txt
//        (PAO model -> HiGHS, second filter, abundance constraint; profile.rs:2689-2882, 1229-1285, 3028-3289).
// All arithmetic on reads/nodes/paths runs in libpantax_gpu.so; this file does file I/O, the f64 tail and TSVs.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cmath>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/pantax_gpu.h"

namespace {

struct Options {
    std::string db, gaf, wd = ".", report, range_file, len_file, designated, reads_binning, filter_gaf;
    bool species = false, strain = false, filtered = true, long_read = false, shift = false, force = false;
    double min_species_abundance = 1e-4, fr = -1, min_depth = 0;
    bool host_gfa = false;  // --host-gfa: parse species GFAs with the C++ reader instead of on the device
    int mode = 2, device = 0, chunk_mb = 0;  // chunk_mb: size of the pinned GAF chunks (0: 64 MB)
};

[[noreturn]] void die(const std::string& m) {
    fprintf(stderr, "pantax-gpu-profile: %s\n", m.c_str());
    exit(1);
}
void ck(ptx_ctx* ctx, int rc, const char* what) {
    if (rc != PTX_OK) die(std::string(what) + ": " + (ctx ? ptx_last_error(ctx) : "error") + " (code " + std::to_string(rc) + ")");
}
bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
std::vector<std::string> split(const std::string& s, char d) {
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        size_t b = s.find(d, a);
        out.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return out;
}
// shortest round-trip decimal like polars' CSV writer (ryu); whole numbers keep a ".0"
std::string fmt_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    std::string s(buf, r.ptr);
    if (s.find_first_of(".en") == std::string::npos) s += ".0";
    return s;
}

struct Range { std::string taxid; int64_t start, end; int is_pan; };

std::vector<Range> read_ranges(const std::string& path) {
    std::ifstream f(path);
    if (!f) die("cannot open species range file " + path);
    std::vector<Range> out;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        auto p = split(line, '\t');
        if (p.size() < 3) die("species range file: expected taxid<TAB>start<TAB>end[<TAB>is_pan]: " + line);  // sort_range.rs:15
        out.push_back({p[0], std::stoll(p[1]), std::stoll(p[2]), p.size() > 3 ? std::stoi(p[3]) : 1});
    }
    return out;
}

struct Graph {  // types.rs:51-55
    std::vector<int64_t> nodes_len;
    std::map<std::string, std::vector<uint64_t>> paths;  // BTreeMap: name order
};

// ---- .bin / .bin.lz4 / .bin.zst (zip.rs:236-262) --------------------------------------------------------------------
// The reference wraps the same bincode stream in an LZ4 *frame* (lz4_flex FrameEncoder, zip.rs:192-205) or a zstd frame
// (zstd::Encoder, zip.rs:206-219).  liblz4 / libzstd are resolved with dlopen (the image ships the runtime libraries without
// their headers); the few prototypes used are declared here.
bool slurp(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    if (n) f.read(reinterpret_cast<char*>(out.data()), n);
    return (bool)f;
}
bool lz4_frame_decode(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    static void* h = dlopen("liblz4.so.1", RTLD_NOW);
    if (!h) die("liblz4.so.1 not found: cannot read .bin.lz4 graphs");
    typedef size_t (*create_t)(void**, unsigned);
    typedef size_t (*free_t)(void*);
    typedef size_t (*dec_t)(void*, void*, size_t*, const void*, size_t*, const void*);
    typedef unsigned (*iserr_t)(size_t);
    static create_t create = (create_t)dlsym(h, "LZ4F_createDecompressionContext");
    static free_t release = (free_t)dlsym(h, "LZ4F_freeDecompressionContext");
    static dec_t dec = (dec_t)dlsym(h, "LZ4F_decompress");
    static iserr_t iserr = (iserr_t)dlsym(h, "LZ4F_isError");
    if (!create || !release || !dec || !iserr) die("liblz4.so.1 lacks the LZ4F frame API");
    void* ctx = nullptr;
    if (iserr(create(&ctx, 100 /* LZ4F_VERSION */))) return false;
    out.clear();
    std::vector<uint8_t> buf(4u << 20);
    size_t ip = 0;
    bool ok = true;
    while (ip < in.size()) {
        size_t dn = buf.size(), sn = in.size() - ip;
        const size_t r = dec(ctx, buf.data(), &dn, in.data() + ip, &sn, nullptr);
        if (iserr(r)) { ok = false; break; }
        out.insert(out.end(), buf.begin(), buf.begin() + (std::ptrdiff_t)dn);
        ip += sn;
        if (r == 0 && sn == 0 && dn == 0) break;  // frame complete
    }
    release(ctx);
    return ok;
}
bool zstd_decode(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    static void* h = dlopen("libzstd.so.1", RTLD_NOW);
    if (!h) die("libzstd.so.1 not found: cannot read .bin.zst graphs");
    struct InBuf { const void* src; size_t size, pos; };
    struct OutBuf { void* dst; size_t size, pos; };
    typedef void* (*create_t)();
    typedef size_t (*free_t)(void*);
    typedef size_t (*dec_t)(void*, OutBuf*, InBuf*);
    typedef unsigned (*iserr_t)(size_t);
    static create_t create = (create_t)dlsym(h, "ZSTD_createDStream");
    static free_t release = (free_t)dlsym(h, "ZSTD_freeDStream");
    static dec_t dec = (dec_t)dlsym(h, "ZSTD_decompressStream");
    static iserr_t iserr = (iserr_t)dlsym(h, "ZSTD_isError");
    if (!create || !release || !dec || !iserr) die("libzstd.so.1 lacks the streaming API");
    void* ds = create();
    if (!ds) return false;
    out.clear();
    std::vector<uint8_t> buf(4u << 20);
    InBuf ib{in.data(), in.size(), 0};
    bool ok = true;
    while (ib.pos < ib.size) {
        OutBuf ob{buf.data(), buf.size(), 0};
        const size_t r = dec(ds, &ob, &ib);
        if (iserr(r)) { ok = false; break; }
        out.insert(out.end(), buf.begin(), buf.begin() + (std::ptrdiff_t)ob.pos);
    }
    release(ds);
    return ok;
}

// bincode 1.3 default options: little-endian, fixed-width ints, u64 lengths (zip.rs:185 serialize_into):
//   u64 n; i64 nodes_len[n]; u64 n_paths; repeat { u64 klen; u8 key[klen]; u64 plen; u64 node[plen] }, keys ascending (BTreeMap)
bool parse_bin_graph(const std::vector<uint8_t>& b, Graph& g) {
    size_t p = 0;
    auto rd64 = [&](uint64_t& v) { if (p + 8 > b.size()) return false; memcpy(&v, b.data() + p, 8); p += 8; return true; };
    uint64_t n = 0, np = 0;
    if (!rd64(n) || n > (b.size() - p) / 8) return false;
    g.nodes_len.resize(n);
    memcpy(g.nodes_len.data(), b.data() + p, n * 8);
    p += n * 8;
    if (!rd64(np)) return false;
    for (uint64_t i = 0; i < np; ++i) {
        uint64_t kl = 0, pl = 0;
        if (!rd64(kl) || kl > b.size() - p) return false;
        std::string key(reinterpret_cast<const char*>(b.data() + p), kl);
        p += kl;
        if (!rd64(pl) || pl > (b.size() - p) / 8) return false;
        std::vector<uint64_t> path(pl);
        memcpy(path.data(), b.data() + p, pl * 8);
        p += pl * 8;
        g.paths[key] = std::move(path);
    }
    return p == b.size();
}
// kind: 0 = .bin, 1 = .bin.lz4, 2 = .bin.zst
bool read_bin_graph(const std::string& path, Graph& g, int kind = 0) {
    std::vector<uint8_t> raw, plain;
    if (!slurp(path, raw)) return false;
    if (kind == 1) { if (!lz4_frame_decode(raw, plain)) return false; }
    else if (kind == 2) { if (!zstd_decode(raw, plain)) return false; }
    else plain.swap(raw);
    return parse_bin_graph(plain, g);
}

// The plain bincode layout of zip.rs:185 / types.rs:51-55 (u64 n; i64 nodes_len[n]; u64 n_paths; per path in key order: u64 klen, key,
// u64 plen, u64 node[plen]) - written to <wd>/strain_graphs for the strain tail (strain_tail.py) when the species' graph did not come from a plain .bin.
bool write_bin_graph(const std::string& path, const Graph& g) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = true;
    const auto w64 = [&](uint64_t v) { ok = ok && fwrite(&v, 8, 1, f) == 1; };
    w64((uint64_t)g.nodes_len.size());
    if (!g.nodes_len.empty()) ok = ok && fwrite(g.nodes_len.data(), 8, g.nodes_len.size(), f) == g.nodes_len.size();
    w64((uint64_t)g.paths.size());
    for (auto& kv : g.paths) {
        w64((uint64_t)kv.first.size());
        if (!kv.first.empty()) ok = ok && fwrite(kv.first.data(), 1, kv.first.size(), f) == kv.first.size();
        w64((uint64_t)kv.second.size());
        if (!kv.second.empty()) ok = ok && fwrite(kv.second.data(), 8, kv.second.size(), f) == kv.second.size();
    }
    return fclose(f) == 0 && ok;
}

template <class F>
void for_digit_runs(const std::string& s, bool allow_minus, F fn) {
    size_t i = 0;
    while (i < s.size()) {
        if (isdigit((unsigned char)s[i])) {
            bool neg = allow_minus && i > 0 && s[i - 1] == '-';
            int64_t v = 0;
            while (i < s.size() && isdigit((unsigned char)s[i])) v = v * 10 + (s[i++] - '0');
            fn(neg ? -v : v);
        } else {
            ++i;
        }
    }
}

// profile.rs:466-545 (previous = 0)
// `zip` (zip.rs:78-171 read_and_zip_gfa, what the reference's database step writes into the .bin): a W line whose walk starts
// with '<' and a P line whose first step ends with '-' are stored reversed; min / max node id over all path lines come back 1-based.
struct ZipInfo { bool on = false; uint64_t min1 = ~0ull, max1 = 0; bool any = false; };
bool read_gfa_graph(const std::string& path, Graph& g, ZipInfo* zip = nullptr) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    size_t idx = 0;
    const auto space = [](char c) { return c == ' ' || (c >= 9 && c <= 13); };  // what str::trim removes (ASCII)
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        if (line[0] == 'S' || line[0] == 'W' || line[0] == 'P')
            while (!line.empty() && space(line.back())) line.pop_back();  // line.trim() (profile.rs:483, 497); nothing to trim in front
        if (line[0] == 'S') {
            auto p = split(line, '\t');
            if (p.size() < 3) continue;
            size_t id = std::stoull(p[1]) - 1;
            if (id != idx) die("Node ID out of order or mismatch (profile.rs:489) in " + path);
            ++idx;
            if (p[2].empty()) die("Node length 0 appears in the GFA (profile.rs:494): " + path);
            g.nodes_len.push_back((int64_t)p[2].size());
        } else if (line[0] == 'W' || line[0] == 'P') {
            auto p = split(line, '\t');
            std::string hap;
            std::vector<uint64_t> nodes;
            if (p[0] == "W") {
                hap = p.size() > 1 ? p[1] : "";
                for_digit_runs(p.back(), true, [&](int64_t v) { nodes.push_back((uint64_t)(v - 1)); });
            } else {
                hap = p.size() > 1 ? split(p[1], '#')[0] : "";
                for_digit_runs(p.size() > 2 ? p[2] : "", false, [&](int64_t v) { nodes.push_back((uint64_t)(v - 1)); });
            }
            if (zip && zip->on) {
                const std::string& fld = p[0] == "W" ? p.back() : (p.size() > 2 ? p[2] : std::string());
                const bool rev = p[0] == "W" ? (!fld.empty() && fld[0] == '<') : [&] { const std::string first = fld.substr(0, fld.find(',')); return !first.empty() && first.back() == '-'; }();
                if (rev) std::reverse(nodes.begin(), nodes.end());  // zip.rs:124, 137, 147-149
                if (nodes.empty()) die("path line without nodes in " + path + " (zip.rs:151 unwraps the minimum)");
                for (uint64_t v : nodes) { zip->min1 = std::min(zip->min1, v + 1); zip->max1 = std::max(zip->max1, v + 1); }
                zip->any = true;
            }
            auto& dst = g.paths[hap];  // same hap id: chromosomes are concatenated (profile.rs:540)
            dst.insert(dst.end(), nodes.begin(), nodes.end());
        }
    }
    return true;
}

// profile.rs:1028-1051
std::vector<double> zscore_filter(const std::vector<double>& d, double thr) {
    if (d.empty()) return {};
    double mean = 0;
    for (double x : d) mean += x;
    mean /= (double)d.size();
    double var = 0;
    for (double x : d) var += (x - mean) * (x - mean);
    double sd = std::sqrt(var / (double)d.size());
    if (sd == 0.0) return {};
    std::vector<double> out;
    for (double x : d)
        if (std::fabs((x - mean) / sd) < thr) out.push_back(x);
    return out;
}
double round2(double x) { return std::round(x * 100.0) / 100.0; }

void usage() {
    puts("pantax-gpu-profile --db DIR --gaf FILE|- [--wd DIR] [--species] [--strain] [-R reads_classification.tsv]\n"
         "                   [--reads-binning reads_classification.tsv]   (with --strain only: species column of the GAF rows)\n"
         "                   [-a MIN_SPECIES_ABUND=1e-4] [--fr F] [--long-read] [--shift] [--no-filter] [--smode 0|1|2]\n"
         "                   [--ds TAXID,TAXID] [--range-file F] [--len-file F] [--min-depth D] [--device N] [--chunk-mb M]\n"
         "                   [--force]   (run a stage although its table exists in --wd; default: skip it as the reference does)\n"
         "                   --filter-gaf FILE [--device N]   (long reads: FILE -> <stem>_filtered.gaf, gaf_filter.rs:44-97)\n"
         "                   --dump-graph FILE | --convert-graph FILE OUT.bin | --zip-gfa FILE.gfa OUTDIR RANGE_FILE   (graph files, no GPU)\n"
         "GPU implementation of PanTax's profiling stage (read classification, species abundance, node coverage and\n"
         "strain statistics).  Needs a CUDA device; there is no CPU fallback.");
}

}  // namespace

int run(int argc, char** argv);
int main(int argc, char** argv) {
    // a malformed number in an option or an input table (std::stod / stoull / ...) ends the run with a message, not with an abort
    try {
        return run(argc, argv);
    } catch (const std::exception& e) {
        fprintf(stderr, "pantax-gpu-profile: malformed number or allocation failure (%s)\n", e.what());
        return 1;
    }
}
int run(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) die("missing value for " + a); return argv[++i]; };
        if (a == "--db") o.db = next();
        else if (a == "--gaf") o.gaf = next();
        else if (a == "--wd" || a == "-T") o.wd = next();
        else if (a == "--species") o.species = true;
        else if (a == "--strain") o.strain = true;
        else if (a == "-R" || a == "--report") o.report = next();
        else if (a == "-a") o.min_species_abundance = std::stod(next());
        else if (a == "--reads-binning") o.reads_binning = next();
        else if (a == "--fr") o.fr = std::stod(next());
        else if (a == "--long-read") o.long_read = true;
        else if (a == "--shift") o.shift = true;
        else if (a == "--no-filter") o.filtered = false;
        else if (a == "--smode") o.mode = std::stoi(next());
        else if (a == "--ds") o.designated = next();
        else if (a == "--range-file") o.range_file = next();
        else if (a == "--len-file") o.len_file = next();
        else if (a == "--min-depth") o.min_depth = std::stod(next());
        else if (a == "--device") o.device = std::stoi(next());
        else if (a == "--chunk-mb") o.chunk_mb = std::stoi(next());
        else if (a == "--host-gfa") o.host_gfa = true;
        else if (a == "--filter-gaf") o.filter_gaf = next();
        else if (a == "--time-gfa") {
            // graph load of one species GFA both ways: the C++ reader on the host, and ptx_upload_graph_gfa (H2D + kernels + compact arrays back)
            const std::string path = next();
            std::vector<uint8_t> bytes;
            if (!slurp(path, bytes) || bytes.empty()) die("cannot read " + path);
            auto t0 = std::chrono::steady_clock::now();
            Graph g;
            if (!read_gfa_graph(path, g)) die("cannot parse " + path);
            const double t_host = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            ptx_ctx* c = nullptr;
            if (ptx_create(o.device, &c) != PTX_OK) die("no usable CUDA device");
            const char* nm = "1";
            const int64_t st1 = 1, en1 = (int64_t)g.nodes_len.size();
            ck(c, ptx_set_ranges(c, 1, &nm, &st1, &en1), "ptx_set_ranges");
            double t_dev = 0;
            for (int rep = 0; rep < 2; ++rep) {  // the second call has the context warm
                t0 = std::chrono::steady_clock::now();
                ck(c, ptx_upload_graph_gfa(c, 0, bytes.data(), bytes.size()), "ptx_upload_graph_gfa");
                t_dev = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            }
            size_t steps = 0;
            for (auto& kv : g.paths) steps += kv.second.size();
            if ((size_t)ptx_species_path_steps(c, 0) != steps || ptx_species_paths(c, 0) != (int64_t)g.paths.size()) die("device and host parse disagree");
            printf("{\"gfa_bytes\": %zu, \"nodes\": %zu, \"paths\": %zu, \"path_steps\": %zu, \"host_reader_s\": %.4f, \"device_parse_s\": %.4f}\n", bytes.size(),
                   g.nodes_len.size(), g.paths.size(), steps, t_host, t_dev);
            ptx_destroy(c);
            return 0;
        }
        else if (a == "--force") o.force = true;
        else if (a == "--dump-graph") {
            // reader check without a GPU: parse one graph file (.bin / .bin.lz4 / .bin.zst / .gfa by its extension) and print it
            const std::string path = next();
            Graph g;
            const auto ends = [&](const char* e) { const size_t n = strlen(e); return path.size() >= n && path.compare(path.size() - n, n, e) == 0; };
            const bool ok = ends(".lz4") ? read_bin_graph(path, g, 1) : ends(".zst") ? read_bin_graph(path, g, 2) : ends(".bin") ? read_bin_graph(path, g, 0)
                                                                                                                   : read_gfa_graph(path, g);
            if (!ok) die("cannot read graph " + path);
            printf("nodes %zu\n", g.nodes_len.size());
            for (size_t i = 0; i < g.nodes_len.size(); ++i) printf("%lld%c", (long long)g.nodes_len[i], i + 1 == g.nodes_len.size() ? '\n' : ' ');
            for (auto& kv : g.paths) {
                printf("path %s %zu\n", kv.first.c_str(), kv.second.size());
                for (size_t i = 0; i < kv.second.size(); ++i) printf("%llu%c", (unsigned long long)kv.second[i], i + 1 == kv.second.size() ? '\n' : ' ');
            }
            return 0;
        }
        else if (a == "--convert-graph") {
            // no GPU: any graph file (.bin / .bin.lz4 / .bin.zst / .gfa by its extension) -> plain .bin (what <wd>/strain_graphs/<taxid>.bin is)
            const std::string path = next(), out = next();
            Graph g;
            const auto ends = [&](const char* e) { const size_t n = strlen(e); return path.size() >= n && path.compare(path.size() - n, n, e) == 0; };
            const bool ok = ends(".lz4") ? read_bin_graph(path, g, 1) : ends(".zst") ? read_bin_graph(path, g, 2) : ends(".bin") ? read_bin_graph(path, g, 0)
                                                                                                                   : read_gfa_graph(path, g);
            if (!ok) die("cannot read graph " + path);
            if (!write_bin_graph(out, g)) die("cannot write " + out);
            return 0;
        }
        else if (a == "--zip-gfa") {
            // no GPU: zip::zip (zip.rs:316-327) for one species GFA - <outdir>/<stem>.bin in the reference's bincode layout (kept if it
            // exists) and the row `stem \t min \t max \t is_pan` appended to the range file
            const std::string path = next(), outdir = next(), range_out = next();
            Graph g;
            ZipInfo z;
            z.on = true;
            if (!read_gfa_graph(path, g, &z)) die("cannot read graph " + path);
            if (!z.any) die("no path lines in " + path + " (zip.rs:159 unwraps the minimum)");
            std::string stem = path.substr(path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
            if (stem.find('.') != std::string::npos && stem.find_last_of('.') > 0) stem = stem.substr(0, stem.find_last_of('.'));
            const std::string out = outdir + "/" + stem + ".bin";
            if (!exists(out) && !write_bin_graph(out, g)) die("cannot write " + out);
            FILE* rf = fopen(range_out.c_str(), "ab");
            if (!rf) die("cannot append to " + range_out);
            fprintf(rf, "%s\t%llu\t%llu\t%d\n", stem.c_str(), (unsigned long long)z.min1, (unsigned long long)z.max1, g.paths.size() > 1 ? 1 : 0);
            fclose(rf);
            return 0;
        }
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else die("unknown option " + a);
    }
    if (!o.filter_gaf.empty()) {
        // gaf_filter::filter_max_alignment_mt (gaf_filter.rs:44-97) as alignment.rs:171 uses it: <dir>/<stem>_filtered.gaf holds, per read
        // id, the first line that is its best alignment (max matches, then identity) with mapq > 20 and a span > 1000
        std::vector<uint8_t> bytes;
        if (!slurp(o.filter_gaf, bytes)) die("cannot read " + o.filter_gaf);
        ptx_ctx* fc = nullptr;
        if (ptx_create(o.device, &fc) != PTX_OK) die("no usable CUDA device (this tool has no CPU fallback)");
        int64_t cap = 1, n_out = 0;
        for (uint8_t c : bytes) cap += c == '\n';
        std::vector<uint64_t> off((size_t)cap);
        if (!bytes.empty()) ck(fc, ptx_filter_gaf(fc, bytes.data(), bytes.size(), off.data(), cap, &n_out), "ptx_filter_gaf");
        const size_t slash = o.filter_gaf.find_last_of('/');
        const std::string dir = slash == std::string::npos ? "" : o.filter_gaf.substr(0, slash + 1);
        std::string stem = slash == std::string::npos ? o.filter_gaf : o.filter_gaf.substr(slash + 1);
        if (stem.find_last_of('.') != std::string::npos && stem.find_last_of('.') > 0) stem = stem.substr(0, stem.find_last_of('.'));
        const std::string out = dir + stem + "_filtered.gaf";
        FILE* f = fopen(out.c_str(), "wb");
        if (!f) die("cannot write " + out);
        for (int64_t k = 0; k < n_out; ++k) {
            const size_t b = (size_t)off[(size_t)k];
            const void* nl = memchr(bytes.data() + b, '\n', bytes.size() - b);
            size_t e = nl ? (size_t)((const uint8_t*)nl - bytes.data()) : bytes.size();
            if (e > b && bytes[e - 1] == '\r') --e;  // BufRead::lines drops "\r\n"
            fwrite(bytes.data() + b, 1, e - b, f);
            fputc('\n', f);
        }
        fclose(f);
        printf("Filtered GAF file written to: %s\n", out.c_str());  // gaf_filter.rs:95
        ptx_destroy(fc);
        return 0;
    }
    if (!o.species && !o.strain) die("Please choose profiling level with --species or/and --strain.");  // profile.rs:73-75
    if (o.db.empty() || o.gaf.empty()) { usage(); return 1; }
    if (!o.force) {
        // profile.rs:3333-3425: a stage whose table exists is not run again - `--species` is skipped when species_abundance.txt exists
        // (a `--strain` request then resumes from reads_classification.tsv, :3365), `--strain` when strain_abundance.txt exists
        const bool sp_done = exists(o.wd + "/species_abundance.txt"), st_done = exists(o.wd + "/strain_abundance.txt");
        if (o.species && !sp_done) {
            if (o.strain && st_done) o.strain = false;  // :3361
        } else if (o.strain && !st_done) {
            o.species = false;
        } else {
            fprintf(stderr, "%s\n", o.species && o.strain ? "Species and strain profiling abundance files both exist." :
                                    o.species ? "Species profiling abundance file exists." : "Strain profiling abundance file exists.");
            return 0;  // :3414-3422 (--force runs the stages anyway)
        }
    }
    if (o.fr < 0) o.fr = o.long_read ? 0.5 : 0.3;  // main.rs:107-113
    if (o.range_file.empty()) o.range_file = o.db + "/species_range.txt";
    if (o.len_file.empty()) o.len_file = o.db + "/species_genomes_stats.txt";
    mkdir(o.wd.c_str(), 0755);

    const std::vector<Range> ranges = read_ranges(o.range_file);
    if (ranges.empty()) die("species range file is empty");
    ptx_ctx* ctx = nullptr;
    if (ptx_create(o.device, &ctx) != PTX_OK) die("no usable CUDA device (this tool has no CPU fallback)");
    {
        std::vector<const char*> names;
        std::vector<int64_t> st, en;
        for (auto& r : ranges) { names.push_back(r.taxid.c_str()); st.push_back(r.start); en.push_back(r.end); }
        ck(ctx, ptx_set_ranges(ctx, (int)ranges.size(), names.data(), st.data(), en.data()), "ptx_set_ranges");
    }

    // ---- strain-only resume (profile.rs:3365-3385): the species column comes from the reads binning file
    // (<wd>/reads_classification.tsv unless given, profile.rs:179-182), row-aligned with the GAF
    if (o.strain && !o.species) {
        std::string rb = !o.reads_binning.empty() && exists(o.reads_binning) ? o.reads_binning : o.wd + "/reads_classification.tsv";
        if (!exists(rb)) die("Neither reads binning file '" + o.reads_binning + "' nor '" + o.wd + "/reads_classification.tsv' is a valid file path");
        std::map<std::string, uint32_t> idx;
        for (size_t i = 0; i < ranges.size(); ++i) idx.emplace(ranges[i].taxid, (uint32_t)i);  // first row wins, as in rcls.rs:253
        std::ifstream f(rb);
        std::string line;
        std::vector<uint32_t> labels;
        while (std::getline(f, line)) {
            auto p = split(line, '\t');
            if (p.size() < 3) die("reads binning file: a row has fewer than 3 columns: " + rb);
            auto it = idx.find(p[2]);
            if (it == idx.end() && p[2] != "U") die("reads binning file: species '" + p[2] + "' is not in the range file");
            labels.push_back(it == idx.end() ? PTX_LABEL_UNCLASSIFIED : it->second);
        }
        ck(ctx, ptx_ingest_labels(ctx, labels.data(), (int64_t)labels.size()), "ptx_ingest_labels");
        fprintf(stderr, "- Species column of %zu rows taken from %s\n", labels.size(), rb.c_str());
    }

    // ---- read classification + species counts: stream the GAF through the library in 256 MB host chunks
    // ("-": the aligner's stdout, e.g. `vg giraffe -o gaf ... | pantax-gpu-profile --gaf -`, alignment.rs:18-26)
    FILE* gf = o.gaf == "-" ? stdin : fopen(o.gaf.c_str(), "rb");
    if (!gf) die("cannot open GAF mapping file " + o.gaf);
    // Double-buffered streaming: a reader thread fills pinned chunks (a pipe returns short reads: it fills each chunk) while this
    // thread hands the previous one to the library, whose H2D copy and kernels run asynchronously - read(), PCIe and the GPU overlap.
    const size_t CH = (size_t)(o.chunk_mb > 0 ? o.chunk_mb : 64) << 20;
    constexpr int NSLOT = 3;
    struct Slot { void* pin = nullptr; size_t n = 0; bool last = false; };
    Slot slots[NSLOT];
    for (auto& sl : slots)
        if (ptx_host_alloc(CH, &sl.pin) != PTX_OK) die("pinned allocation failed");
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> ready, freeq;
    for (int i = 0; i < NSLOT; ++i) freeq.push_back(i);
    const auto t_stream0 = std::chrono::steady_clock::now();
    std::thread reader([&] {
        for (;;) {
            int si;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !freeq.empty(); });
                si = freeq.front();
                freeq.pop_front();
            }
            Slot& sl = slots[si];
            sl.n = 0;
            while (sl.n < CH) { size_t k = fread((char*)sl.pin + sl.n, 1, CH - sl.n, gf); if (k == 0) break; sl.n += k; }
            sl.last = sl.n < CH;
            {
                std::lock_guard<std::mutex> lk(mu);
                ready.push_back(si);
            }
            cv.notify_all();
            if (sl.last) return;
        }
    });
    // reads_classification.tsv (profile.rs:3337-3351) needs read id, mapq and read length of every row: they are cut out of each chunk
    // while the next one is being read (about 35 bytes per row are kept, not the text)
    const bool want_report = !o.report.empty();
    std::string rep_rows, rep_carry;  // "id\tmapq\trlen\n" per GAF row; the unterminated tail of the previous chunk
    auto is_int = [](const char* b, size_t n) {
        size_t k = (n && (b[0] == '+' || b[0] == '-')) ? 1 : 0;
        if (k == n || n - k > 18) return false;
        for (; k < n; ++k) if (!isdigit((unsigned char)b[k])) return false;
        return true;
    };
    auto report_line = [&](const char* b, size_t l) {
        if (l && b[l - 1] == '\r') --l;
        if (!l || b[0] == '@') return;
        const char* f[12]; size_t fl[12]; int nf = 0;
        size_t st = 0;
        for (size_t i = 0; i <= l && nf < 12; ++i)
            if (i == l || b[i] == '\t') { f[nf] = b + st; fl[nf] = i - st; ++nf; st = i + 1; }
        // polars' CsvWriter (rcls.rs:415-418, QuoteStyle::Necessary): a string holding the quote character or a line break is
        // written between quotes with its quotes doubled (a read id cannot hold the separator)
        if (memchr(f[0], '"', fl[0]) || memchr(f[0], '\r', fl[0])) {
            rep_rows.push_back('"');
            for (size_t k = 0; k < fl[0]; ++k) { if (f[0][k] == '"') rep_rows.push_back('"'); rep_rows.push_back(f[0][k]); }
            rep_rows.push_back('"');
        } else {
            rep_rows.append(f[0], fl[0]);
        }
        rep_rows.push_back('\t');
        if (nf > 11 && is_int(f[11], fl[11])) rep_rows.append(f[11], fl[11]);
        rep_rows.push_back('\t');
        if (nf > 1 && is_int(f[1], fl[1])) rep_rows.append(f[1], fl[1]);
        rep_rows.push_back('\n');
    };
    size_t gaf_bytes = 0;
    for (;;) {
        int si;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !ready.empty(); });
            si = ready.front();
            ready.pop_front();
        }
        Slot& sl = slots[si];
        gaf_bytes += sl.n;
        ck(ctx, ptx_ingest_gaf(ctx, (const uint8_t*)sl.pin, sl.n, sl.last ? 1 : 0), "ptx_ingest_gaf");  // returns when the chunk is on the device
        if (want_report) {
            const char* b = (const char*)sl.pin;
            size_t i = 0;
            if (!rep_carry.empty()) {
                const void* nl = memchr(b, '\n', sl.n);
                const size_t e = nl ? (size_t)((const char*)nl - b) : sl.n;
                rep_carry.append(b, e);
                if (nl || sl.last) { report_line(rep_carry.data(), rep_carry.size()); rep_carry.clear(); }
                i = nl ? e + 1 : sl.n;
            }
            while (i < sl.n) {
                const void* nl = memchr(b + i, '\n', sl.n - i);
                if (!nl && !sl.last) { rep_carry.assign(b + i, sl.n - i); break; }
                const size_t e = nl ? (size_t)((const char*)nl - b) : sl.n;
                report_line(b + i, e - i);
                i = e + 1;
            }
        }
        const bool last = sl.last;
        {
            std::lock_guard<std::mutex> lk(mu);
            freeq.push_back(si);
        }
        cv.notify_all();
        if (last) break;
    }
    reader.join();
    if (gf != stdin) fclose(gf);
    const double t_stream = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_stream0).count();
    ck(ctx, ptx_finalize(ctx), "ptx_finalize");
    for (auto& sl : slots) ptx_host_free(sl.pin);
    const int64_t R = ptx_num_records(ctx);
    const int S = (int)ranges.size();
    std::vector<int64_t> counts((size_t)S * 4);
    ck(ctx, ptx_species_counts(ctx, counts.data()), "ptx_species_counts");
    fprintf(stderr, "- Read classification: %lld GAF records, ids unique: %s\n", (long long)R, ptx_ids_unique(ctx) ? "yes" : "NO (duplicate read ids, profile.rs:460)");
    fprintf(stderr, "- GAF stream: %.1f MB in %.3f s = %.1f MB/s (%s, %d pinned chunks of %zu MB)\n", gaf_bytes / 1e6, t_stream, gaf_bytes / 1e6 / std::max(t_stream, 1e-9),
            o.gaf == "-" ? "stdin" : "file", NSLOT, CH >> 20);

    if (want_report) {  // profile.rs:3337-3351
        std::vector<uint32_t> labels((size_t)std::max<int64_t>(R, 1));
        ck(ctx, ptx_read_labels(ctx, labels.data()), "ptx_read_labels");
        FILE* rf = fopen(o.report.c_str(), "wb");
        if (!rf) die("cannot write " + o.report);
        size_t i = 0;
        int64_t rec = 0;
        while (i < rep_rows.size()) {
            const size_t e = rep_rows.find('\n', i);
            const size_t t2 = rep_rows.rfind('\t', e);  // in front of the read length (ids hold no tab)
            const uint32_t lab = rec < R ? labels[(size_t)rec] : PTX_LABEL_UNCLASSIFIED;
            fwrite(rep_rows.data() + i, 1, t2 + 1 - i, rf);
            fputs(lab == PTX_LABEL_UNCLASSIFIED ? "U" : ranges[lab].taxid.c_str(), rf);
            fputc('\t', rf);
            fwrite(rep_rows.data() + t2 + 1, 1, e - t2, rf);  // read length + '\n'
            ++rec;
            i = e + 1;
        }
        fclose(rf);
    }

    // ---- species abundance (float tail of profile.rs:299-349)
    std::map<std::string, double> species_len;
    {
        std::ifstream f(o.len_file);
        if (!f) die("cannot open species length file " + o.len_file);
        std::string line;
        while (std::getline(f, line)) {
            auto p = split(line, '\t');
            if (p.size() >= 2) species_len[p[0]] = std::stod(p[1]);
        }
    }
    int eq = 0;
    int64_t rl0 = 0;
    ck(ctx, ptx_equal_length(ctx, &eq, &rl0), "ptx_equal_length");
    struct Row { std::string taxid; double rel, abs; };
    std::vector<Row> table;
    double total = 0;
    for (int s = 0; s < S; ++s) {
        const int64_t rc = counts[4 * s], sl = counts[4 * s + 1], lm = counts[4 * s + 2], uq = counts[4 * s + 3];
        if (rc == 0) continue;
        if (o.filtered && !(uq > 0 && (double)lm > (double)rc / 10.0)) continue;  // profile.rs:239-245
        const double base = eq ? (double)(rc * rl0) : (double)sl;                 // :246 / :291
        auto it = species_len.find(ranges[s].taxid);
        const double ab = it == species_len.end() ? NAN : base / it->second;      // :333-337 (left join)
        table.push_back({ranges[s].taxid, 0, ab});
        if (!std::isnan(ab)) total += ab;
    }
    for (auto& r : table) r.rel = r.abs / total;
    std::stable_sort(table.begin(), table.end(), [](const Row& a, const Row& b) { return a.rel > b.rel; });
    if (o.strain && !o.species) {
        // profile.rs:3399-3414: the strain-only run takes the species table from the existing species_abundance.txt
        const std::string sa = o.wd + "/species_abundance.txt";
        std::ifstream f(sa);
        if (!f) die("strain-only run: " + sa + " does not exist (run with --species first)");
        table.clear();
        std::string line;
        std::getline(f, line);  // header
        while (std::getline(f, line)) {
            auto p = split(line, '\t');
            if (p.size() >= 3) table.push_back({p[0], std::stod(p[1]), std::stod(p[2])});
        }
    } else {
        FILE* f = fopen((o.wd + "/species_abundance.txt").c_str(), "wb");
        if (!f) die("cannot write species_abundance.txt");
        fprintf(f, "species_taxid\tpredicted_abundance\tpredicted_coverage\n");
        for (auto& r : table) fprintf(f, "%s\t%s\t%s\n", r.taxid.c_str(), fmt_f64(r.rel).c_str(), fmt_f64(r.abs).c_str());
        fclose(f);
    }
    fprintf(stderr, "- Species level profiling: %zu species\n", table.size());
    if (!o.strain) { ptx_destroy(ctx); return 0; }

    // ---- strain level: graphs of the species that pass load_species_range (profile.rs:553-656)
    std::set<std::string> wanted;
    if (!o.designated.empty() && o.designated != "None")
        for (auto t : split(o.designated, ',')) {  // profile.rs:581-585: pieces are trimmed, empty ones dropped
            while (!t.empty() && isspace((unsigned char)t.back())) t.pop_back();
            while (!t.empty() && isspace((unsigned char)t.front())) t.erase(t.begin());
            if (!t.empty()) wanted.insert(t);
        }
    std::map<std::string, double> rel_of;
    for (auto& r : table) rel_of[r.taxid] = r.rel;
    std::vector<int> chosen;
    for (int s = 0; s < S; ++s) {
        if (o.mode == 0 && ranges[s].is_pan != 0) continue;
        if (o.mode == 1 && ranges[s].is_pan != 1) continue;
        if (!wanted.empty() && !wanted.count(ranges[s].taxid)) continue;
        auto it = rel_of.find(ranges[s].taxid);
        if (it == rel_of.end() || !(it->second > o.min_species_abundance)) continue;
        chosen.push_back(s);
    }
    std::map<int, Graph> graphs;
    std::set<int> not_plain;  // species whose Graph was decoded from .bin.lz4 / .bin.zst / GFA: the strain tail gets it as a plain .bin
    for (int s : chosen) {
        Graph g;
        // profile.rs:2888-2932: <db>/species_graph_info/<taxid>.bin | .bin.lz4 | .bin.zst (zip.rs:236-262), else <db>/species_gfa/<taxid>.gfa
        const std::string bin = o.db + "/species_graph_info/" + ranges[s].taxid + ".bin";
        const std::string gfa = o.db + "/species_gfa/" + ranges[s].taxid + ".gfa";
        const bool plain_bin = exists(bin) && read_bin_graph(bin, g, 0);
        const bool have_bin = plain_bin || (exists(bin + ".lz4") && read_bin_graph(bin + ".lz4", g, 1)) ||
                              (exists(bin + ".zst") && read_bin_graph(bin + ".zst", g, 2));
        if (!plain_bin) not_plain.insert(s);
        if (have_bin) {
            std::vector<uint64_t> off{0}, flat;
            for (auto& kv : g.paths) { flat.insert(flat.end(), kv.second.begin(), kv.second.end()); off.push_back(flat.size()); }
            if (flat.empty()) flat.push_back(0);
            ck(ctx, ptx_upload_graph(ctx, s, g.nodes_len.data(), (int64_t)g.nodes_len.size(), off.data(), flat.data(), (int64_t)g.paths.size()), "ptx_upload_graph");
        } else if (exists(gfa) && !o.host_gfa) {
            // the GFA text is parsed on the device (read_gfa, profile.rs:466-545); the Graph the tables below need comes back compact
            std::vector<uint8_t> bytes;
            if (!slurp(gfa, bytes) || bytes.empty()) die("cannot read " + gfa);
            ck(ctx, ptx_upload_graph_gfa(ctx, s, bytes.data(), bytes.size()), "ptx_upload_graph_gfa");
            const int64_t n = ptx_species_nodes(ctx, s), H = ptx_species_paths(ctx, s), P = ptx_species_path_steps(ctx, s);
            g.nodes_len.resize((size_t)n);
            std::vector<uint64_t> off((size_t)H + 1), flat((size_t)std::max<int64_t>(P, 1));
            ck(ctx, ptx_species_graph(ctx, s, g.nodes_len.data(), off.data(), flat.data()), "ptx_species_graph");
            for (int64_t h = 0; h < H; ++h) {
                char name[4096];
                if (ptx_species_path_name(ctx, s, h, name, sizeof name) < 0) die("ptx_species_path_name failed");
                g.paths[name].assign(flat.begin() + (ptrdiff_t)off[(size_t)h], flat.begin() + (ptrdiff_t)off[(size_t)h + 1]);
            }
        } else if (exists(gfa) && read_gfa_graph(gfa, g)) {
            std::vector<uint64_t> off{0}, flat;
            for (auto& kv : g.paths) { flat.insert(flat.end(), kv.second.begin(), kv.second.end()); off.push_back(flat.size()); }
            if (flat.empty()) flat.push_back(0);
            ck(ctx, ptx_upload_graph(ctx, s, g.nodes_len.data(), (int64_t)g.nodes_len.size(), off.data(), flat.data(), (int64_t)g.paths.size()), "ptx_upload_graph");
        } else {
            die("gfa information file for " + ranges[s].taxid + " does not exist. Please check database.");  // profile.rs:2929
        }
        graphs[s] = std::move(g);
    }
    if (chosen.empty()) { fprintf(stderr, "The filtering before strain profiling has removed all species.\n"); ptx_destroy(ctx); return 0; }
    ck(ctx, ptx_commit_graphs(ctx), "ptx_commit_graphs");
    ck(ctx, ptx_finalize(ctx), "ptx_finalize (coverage)");  // coverage pass over the record tables kept on the device (no text is parsed again)

    mkdir((o.wd + "/strain_inputs").c_str(), 0755);
    for (int s : chosen) {
        const Graph& g = graphs[s];
        const int64_t n = (int64_t)g.nodes_len.size(), H = (int64_t)g.paths.size(), T = ptx_species_trios(ctx, s);
        std::vector<double> depth((size_t)n), tdepth((size_t)std::max<int64_t>(T, 1));
        std::vector<uint64_t> cov((size_t)n);
        std::vector<uint32_t> owner((size_t)std::max<int64_t>(T, 1));
        std::vector<int64_t> sc((size_t)std::max<int64_t>(H, 1)), sl((size_t)std::max<int64_t>(H, 1)), U((size_t)std::max<int64_t>(H, 1)), nz((size_t)std::max<int64_t>(H, 1));
        int rc = ptx_node_depth(ctx, s, depth.data());
        if (rc == PTX_E_START_GT_LEN) die(std::string("read start is bigger than node len (profile.rs:854) in species ") + ranges[s].taxid);
        ck(ctx, rc, "ptx_node_depth");
        ck(ctx, ptx_node_cov(ctx, s, cov.data()), "ptx_node_cov");
        ck(ctx, ptx_trio_depth(ctx, s, tdepth.data()), "ptx_trio_depth");
        ck(ctx, ptx_trio_table(ctx, s, nullptr, nullptr, owner.data()), "ptx_trio_table");
        ck(ctx, ptx_path_sums(ctx, s, sc.data(), sl.data()), "ptx_path_sums");
        ck(ctx, ptx_hap_trio_counts(ctx, s, U.data(), nz.data()), "ptx_hap_trio_counts");
        {
            FILE* f = fopen((o.wd + "/strain_inputs/" + ranges[s].taxid + ".nodes.tsv").c_str(), "wb");
            fprintf(f, "node\tlen\tdepth\tcovered_bases\n");
            for (int64_t i = 0; i < n; ++i)
                if (depth[(size_t)i] > 0) fprintf(f, "%lld\t%lld\t%s\t%llu\n", (long long)i, (long long)g.nodes_len[(size_t)i], fmt_f64(depth[(size_t)i]).c_str(), (unsigned long long)cov[(size_t)i]);
            fclose(f);
        }
        if (not_plain.count(s)) {
            mkdir((o.wd + "/strain_graphs").c_str(), 0755);
            if (!write_bin_graph(o.wd + "/strain_graphs/" + ranges[s].taxid + ".bin", g)) die("cannot write strain_graphs/" + ranges[s].taxid + ".bin");
        }
        // first_filter_paths (profile.rs:1080-1227)
        std::vector<std::string> hap_names;
        for (auto& kv : g.paths) hap_names.push_back(kv.first);
        FILE* f = fopen((o.wd + "/strain_inputs/" + ranges[s].taxid + ".paths.tsv").c_str(), "wb");
        fprintf(f, "hap_id\tunique_trios\tunique_trios_covered\tunique_trio_fraction\tuniq_trio_cov_mean\tpath_base_cov\tsum_cov\tsum_len\tpossible\n");
        std::vector<std::vector<double>> per_hap((size_t)H);
        if (H > 1 && T > 0) {
            // The f64 sums of zscore_filter / frequencies_mean run over a hap's trios in the REFERENCE's trio numbering
            // (FxHashSet iteration order, profile.rs:659-716, 1123-1146) - ptx_trio_ref_order reproduces it on the host.
            std::vector<uint64_t> poff{0}, pnodes, keys3((size_t)T * 3), order((size_t)T);
            for (auto& kv : g.paths) {
                pnodes.insert(pnodes.end(), kv.second.begin(), kv.second.end());
                poff.push_back((uint64_t)pnodes.size());
            }
            if (pnodes.empty()) pnodes.push_back(0);
            ck(ctx, ptx_trio_table(ctx, s, keys3.data(), nullptr, nullptr), "ptx_trio_table (keys)");
            if (ptx_trio_ref_order(poff.data(), pnodes.data(), H, keys3.data(), T, order.data()) != PTX_OK)
                die(std::string("ptx_trio_ref_order: the graph and the trio table disagree for species ") + ranges[s].taxid);
            for (int64_t i = 0; i < T; ++i) {
                const size_t t = (size_t)order[(size_t)i];
                if (tdepth[t] > 0.0) per_hap[owner[t]].push_back(tdepth[t]);
            }
        }
        bool all_same = true;
        if (H > 1 && T == 0) {
            auto it0 = g.paths.begin();
            for (auto it = std::next(it0); it != g.paths.end(); ++it) all_same = all_same && (it->second == it0->second);
        }
        double nz_mean = 0;
        {
            double sum = 0;
            size_t c = 0;
            for (double d : depth) { double x = d > o.min_depth ? d : 0.0; if (x > 0) { sum += x; ++c; } }  // profile.rs:2941-2944, 1212-1221
            nz_mean = c ? sum / (double)c : 0.0;
        }
        std::vector<uint32_t> stamp((size_t)n, 0u);
        for (int64_t h = 0; h < H; ++h) {
            std::string frac_s = "", mean_s = "";
            bool possible = false;
            if (H != 1 && T != 0) {
                if (U[(size_t)h] > 0) {  // :1119
                    const double frac = (double)nz[(size_t)h] / (double)U[(size_t)h];
                    frac_s = fmt_f64(round2(frac));
                    auto zf = zscore_filter(per_hap[(size_t)h], 3.0);
                    double fm = 0;
                    for (double x : zf) fm += x;
                    fm = zf.empty() ? 0.0 : fm / (double)zf.size();
                    double thr = o.fr;
                    if (o.shift) thr = fm >= 1.0 ? std::min(o.fr + (0.8 - o.fr) * fm / 100.0, 0.8) : o.fr * fm;  // :1148-1157
                    if (!(frac < thr)) { possible = true; mean_s = fmt_f64(fm); }
                }
            } else if (H != 1) {
                if (all_same) { if (h == 0) { possible = true; mean_s = fmt_f64(round2(nz_mean)); } }
                else possible = true;  // :1206-1209
            } else {
                possible = true;
                mean_s = fmt_f64(round2(nz_mean));
            }
            // profile.rs:2714-2728: RowDVector<f32> x 0/1 incidence = a sequential f32 accumulation over the distinct nodes
            // of the path in node-index order (nalgebra gemv) - not the exact integer sums once they pass 2^24
            float acc_c = 0.0f, acc_l = 0.0f;
            {
                auto it = g.paths.begin();
                std::advance(it, h);
                for (uint64_t v : it->second) stamp[(size_t)v] = (uint32_t)h + 1u;
                for (int64_t v = 0; v < n; ++v)
                    if (stamp[(size_t)v] == (uint32_t)h + 1u) { acc_c += (float)cov[(size_t)v]; acc_l += (float)g.nodes_len[(size_t)v]; }
            }
            const float ratio = acc_c / acc_l;
            fprintf(f, "%s\t%lld\t%lld\t%s\t%s\t%s\t%lld\t%lld\t%d\n", hap_names[(size_t)h].c_str(), (long long)U[(size_t)h], (long long)nz[(size_t)h],
                    frac_s.c_str(), mean_s.c_str(), fmt_f64((double)ratio).c_str(), (long long)sc[(size_t)h], (long long)sl[(size_t)h], possible ? 1 : 0);
        }
        fclose(f);
    }
    fprintf(stderr, "- Strain level statistics: %zu species written to %s/strain_inputs/\n", chosen.size(), o.wd.c_str());
    char stats[2048];
    if (ptx_stats_json(ctx, stats, sizeof stats) == PTX_OK) fprintf(stderr, "- device: %s\n", stats);

    ptx_destroy(ctx);
    return 0;
}
