// Context + C ABI (include/pantax_gpu.h) of the B200-native PanTax hot path.
// Host-side orchestration only: buffers, streams, kernel sequencing, getters.  All
// arithmetic on GAF records, nodes, paths and trios happens in ptx_kernels.cu.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <functional>
#include <map>
#include <vector>

#include "../../include/pantax_gpu.h"
#include "ptx_internal.h"
#include "ptx_fxorder.h"

using namespace ptx;

namespace {

// ---- minimal NCCL surface, resolved with dlopen so that single-GPU use has no NCCL dependency
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;  // all ranks of ONE process (ptx_create_multi)
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;  // NCCL >= 2.18; optional
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommSplit = (decltype(CommSplit))dlsym(h, "ncclCommSplit");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        Send = (decltype(Send))dlsym(h, "ncclSend");
        Recv = (decltype(Recv))dlsym(h, "ncclRecv");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && Send && Recv && GroupStart && GroupEnd;
    }
};
NcclApi g_nccl;

struct SpeciesHost {
    std::string taxid;
    int64_t start = 0, end = 0;
    bool uploaded = false;   // host copy present (before commit)
    bool has_graph = false;  // on device
    int64_t n_nodes = 0, n_paths = 0;
    int64_t node_base = -1, hap_base = -1, trio_base = 0, n_trios = 0;
    std::vector<uint32_t> len;
    std::vector<uint64_t> path_off;
    std::vector<uint32_t> path_nodes;  // local ids
    std::vector<std::string> path_names;  // ptx_upload_graph_gfa only: haplotype ids in path order
};

struct Chunk {
    uint8_t* buf = nullptr;  // cudaMalloc'ed: [PRE][text][padding]
    size_t cap = 0;          // text capacity (bytes) the buffer was padded for
    size_t n = 0;            // text bytes
    size_t padded = 0;       // bytes readable from text (text + '\n' padding)
    uint32_t n_micro = 0;    // 4 KB micro-tiles counted by K1 (multiple of 8, covers every ingest tile)
    uint32_t n_tiles = 0;
    uint32_t tile_bytes = MAX_TILE;  // text bytes per ingest CTA
    uint32_t over_bytes = OVER;      // bytes staged behind the tile (k_ingest_s)
    uint32_t* tile_count = nullptr;  // [n_micro]
    uint64_t* tile_base = nullptr;   // [n_micro + 1] exclusive record prefix, last = total
    uint64_t* scan_scratch = nullptr;
    uint32_t* labels = nullptr;
    // record table + CSR walks (k_ingest -> k_apply)
    uint4* meta_b = nullptr;
    longlong2* meta_a = nullptr;
    unsigned long long* hash_lo = nullptr;
    uint32_t* nodes = nullptr;
    unsigned long long* cursors = nullptr;  // 8 x u64: words [0..3] as uint32 = k_ingest cursors (entry, node, rows, abandoned); [4] = line slots (count pass)
    uint32_t* h_cur = nullptr;              // pinned copy of the four uint32 cursors (single-pass chunks)
    cudaEvent_t done = nullptr;             // recorded behind that copy
    uint4* tile_info = nullptr;             // single-pass: per tile {first entry, line slots, GAF rows, 0}
    uint16_t* row_key = nullptr;            // single-pass: row of the entry within its tile
    unsigned long long* chunk_hist = nullptr;  // single-pass: species counts of this chunk (merged if not abandoned)
    uint32_t tile_info_cap = 0;
    size_t hist_cap = 0;
    uint32_t hist_copies = 1;
    bool single_pass = false, pending = false, labels_ready = false;
    int64_t est_records = 0;
    int64_t slots_cap = 0, nodes_cap = 0;
    int64_t n_slots = 0;
    uint32_t tiles_cap = 0;
    int64_t labels_cap = 0;
    int64_t n_records = 0;
    int64_t row_base = 0;  // GAF rows ingested before this chunk
    uint32_t long_mode = 0;
    bool ingested = false;  // classify pass done
    bool covered = false;   // coverage pass done against the current graph
    cudaEvent_t copied = nullptr;
    bool host_text = false;  // the text came through ptx_ingest_gaf (not a caller-owned device buffer): it may be released once the chunk is resolved
};

struct EvPair { cudaEvent_t a, b; };

}  // namespace

struct ptx_ctx {
    int device = 0;
    cudaStream_t st = nullptr, copy_st = nullptr;
    std::string err;
    // ranges
    std::vector<SpeciesHost> sp;
    int64_t* d_rstart = nullptr; int64_t* d_rend = nullptr; int64_t* d_node_base = nullptr; uint32_t* d_order = nullptr;
    uint32_t* d_sstart = nullptr;
    int disjoint = 0;
    // graph
    GraphDev g;
    bool graphs_committed = false;
    double commit_ms = 0;  // wall time of the last ptx_commit_graphs (host concatenation + upload + path marks + trio table)
    // accumulators independent of the graph
    unsigned long long* d_hist = nullptr;
    unsigned long long* d_hist_g = nullptr;  // all-reduced copy (multi-GPU)
    bool cov_reduced = false;               // coverage accumulators already hold the cross-rank sum
    bool hist_reduced = false;              // this ptx_finalize already all-reduced the species counts (with the coverage sums)
    uint32_t* d_flags = nullptr;  // [0] dup, [1] mixed
    uint32_t* d_err = nullptr;    // [S]
    ulonglong2* d_ds = nullptr;
    uint64_t ds_cap = 0;
    int64_t ds_records = 0;  // upper bound of ids inserted
    int64_t reserve_records = 0;
    // node-coverage scatter variant (IngestArgs::scatter_var; profiles/r2_scatter_bakeoff.md): -1 = choose (pick_scatter_variant)
    int scatter_var = -1;
    uint32_t* d_pair_key = nullptr; unsigned long long* d_pair_val = nullptr; uint8_t* d_pair_tmp = nullptr;  // variant 3
    int64_t pair_cap = 0; size_t pair_tmp_cap = 0;
    int force_rows = 0;  // PTX_TILE_ROWS env override (tests exercise every tile size)
    int force_tile = 0;  // PTX_TILE_BYTES env override for single-pass chunks of k_ingest_s (multiple of 512)
    bool keep_text = false;  // PTX_KEEP_TEXT=1: never hand the text of resolved chunks back (debugging)
    bool no_sort = false;    // PTX_NO_SORT=1: k_ingest_s keeps file order inside a tile (measurements)
    int pf_waves = 1;          // PTX_PF_WAVES: L2 prefetch distance of the ingest kernels in waves of resident CTAs (0: off)
    int n_sm = 148;
    uint32_t ds_epoch = 1;   // id-set slots written in another epoch are empty (ds_new_pass); 1..255
    int l2_hints = 1;        // PTX_L2_HINTS: bit 0 = graph arrays evict-last (k_apply 0.969 -> 0.938 ms), bit 1 = GAF text evict-first (no gain), bit 2 = id-set loads evict-first (slower: 1.00 ms); IngestArgs::pol_*
    bool old_short = false;  // PTX_OLD_INGEST=1: round 1's byte-at-a-time short-read kernel (A/B measurements)
    uint32_t long_tile_max = LONG_TILE_MAX;  // PTX_LONG_TILE: largest tile of k_ingest_l in bytes (multiple of 4096; measurements)
    bool long_new = true;    // PTX_LONG_NEW=0: round 1's k_ingest<long> (warp-cooperative walk decode) instead of k_ingest_l (A/B measurements: 1.66 vs 1.40 ms on configs[2])
    int force_long = -1; // PTX_LONG_MODE env override: 0/1 = never/always use the long-line kernel (tests)
    int64_t test_box_cap = 0;  // PTX_TEST_BOX_CAP env: first outbox capacity (tests force the overflow/restart path of the exchange)
    uint64_t* d_total = nullptr;  // scratch scalar
    // records
    std::vector<Chunk> chunks;
    std::vector<uint8_t> carry;
    int64_t total_records = 0;
    // single-pass ingest: line statistics of the chunks seen so far, and records of chunks whose counts are still in flight
    double seen_mean_line = 0, seen_slots_per_row = 1;
    int64_t pending_records = 0;
    bool single_pass_ok = true;  // PTX_NO_SINGLE_PASS=1 forces the count pass for every chunk
    int64_t n_table_allocs = 0;  // (re)allocations of chunk record tables: must stay flat in a steady stream of chunks
    uint32_t* d_labels_in = nullptr;  // ptx_ingest_labels: species label of GAF rows [0, labels_in_n), row = position over all chunks
    int64_t labels_in_n = 0, labels_in_cap = 0;
    bool dirty = false;          // something ingested / committed since the last finalize
    uint32_t h_flags[4] = {0, 0, 0, 0};  // [0] repeated id, [1] mixed-species id group, [2] exchange box overflow
    std::vector<uint32_t> h_err;
    // reuse: chunk buffers released by ptx_reset, and one grow-only scratch arena for ptx_finalize
    std::vector<Chunk> pool;
    // text buffers handed back by resolved chunks (ptx_ingest_gaf path): {buffer, text capacity, its `copied` event}.  The text of a
    // chunk is only needed until its counts are in (an abandoned single-pass chunk is redone from it) and, for the first chunks,
    // by ptx_equal_length - the replay passes run from the record tables.  Device memory per GAF byte drops from ~3x to ~0.6x.
    struct TextBuf { uint8_t* buf; size_t cap; cudaEvent_t copied; };
    std::vector<TextBuf> text_pool;
    std::vector<int64_t> first_rows;   // multi-GPU: read_len of the first 1000 non-U rows over all ranks (first_rows_exchange)
    bool first_rows_global = false;
    int64_t labelled_rows_known = 0;  // non-U rows of the chunks resolved so far (exact-mode chunks count as 0: their text is kept)
    double seen_nodes_per_byte = 0;   // walk nodes per text byte of the last resolved chunk: sizes the CSR buffer of single-pass chunks
    uint8_t* scratch = nullptr;
    size_t scratch_cap = 0, scratch_off = 0;
    // asynchronous id-group exchange (multi-GPU): boxes filled by k_apply, sent on a side stream during the coverage pass
    cudaStream_t xs = nullptr;
    cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;
    ulonglong2 *outbox = nullptr, *inbox = nullptr;
    ulonglong2** d_box_ptr = nullptr;  // [P] device array handed to k_apply (see IngestArgs::box_ptr)
    // peer-memory boxes: every rank exports one inbox of P slices through CUDA IPC; senders store straight into it
    bool p2p = false;
    uint64_t p2p_cap = 0;              // entries per (sender, owner) slice
    ulonglong2* p2p_inbox = nullptr;   // [P][p2p_cap], slice q is written by rank q
    std::vector<void*> p2p_peer;       // opened IPC mappings of the peers' inboxes
    bool p2p_inprocess = false;        // ptx_create_multi: the peers live in this process (peer access, no IPC mappings to close)
    unsigned long long* out_cursor = nullptr;  // [P] box cursors, [P] "an entry was dropped" marker, then exchange scratch
    uint64_t box_cap = 0, inbox_cap = 0;
    std::vector<unsigned long long> box_sent, recv_done;  // per peer: entries already exchanged by an earlier ptx_finalize
    int64_t ds_entries_bound = 0;  // upper bound of the entries in the id set (own records + merged foreign ids + returned mixed ids)
    // timing
    std::vector<EvPair> ev_count, ev_ingest, ev_apply, ev_final;
    // multi-GPU
    ncclComm_t comm = nullptr;
    ncclComm_t comm_x = nullptr;  // second communicator (ncclCommSplit) for the id-box exchange on the side stream; == comm if unavailable
    int n_ranks = 1, rank = 0;
};

namespace {

struct Trace {  // PTX_TRACE=1: wall time of the finalize phases (synchronising; debugging aid only)
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t;
    explicit Trace(cudaStream_t s) : on(getenv("PTX_TRACE") != nullptr), st(s) { if (on) { cudaStreamSynchronize(st); t = std::chrono::steady_clock::now(); } }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[ptx trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

int fail(ptx_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                               \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? PTX_E_NOMEM : PTX_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                           \
    } while (0)

template <class T>
int dalloc(ptx_ctx* ctx, T** p, size_t n, bool zero = true) {
    *p = nullptr;
    if (n == 0) n = 1;
    CU(cudaMalloc((void**)p, n * sizeof(T)));
    if (zero) CU(cudaMemsetAsync(*p, 0, n * sizeof(T), ctx->st));
    return PTX_OK;
}
template <class T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

void ev_begin(ptx_ctx* ctx, std::vector<EvPair>& v) {
    EvPair e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, ctx->st);
    v.push_back(e);
}
void ev_end(ptx_ctx* ctx, std::vector<EvPair>& v) { cudaEventRecord(v.back().b, ctx->st); }
double ev_sum(std::vector<EvPair>& v) {
    double ms = 0;
    for (auto& e : v) {
        float f = 0;
        if (cudaEventElapsedTime(&f, e.a, e.b) == cudaSuccess) ms += f;
    }
    return ms;
}
void ev_clear(std::vector<EvPair>& v) {
    for (auto& e : v) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    v.clear();
}

RangesView ranges_view(const ptx_ctx* ctx) {
    RangesView R;
    R.start = ctx->d_rstart;
    R.end = ctx->d_rend;
    R.node_base = ctx->d_node_base;
    R.order = ctx->d_order;
    R.sstart = ctx->d_sstart;
    R.S = (int)ctx->sp.size();
    R.disjoint = ctx->disjoint;
    return R;
}

uint32_t log2_ceil(uint64_t v) {
    uint32_t l = 0;
    while ((1ull << l) < v) ++l;
    return l;
}

// grow (and rehash) the id set so that `entries` fit at load <= 1/2
int ds_ensure_total(ptx_ctx* ctx, int64_t entries) {
    const int64_t need = std::max<int64_t>(entries, ctx->reserve_records);
    uint64_t want = 1ull << std::max<uint32_t>(16, log2_ceil((uint64_t)need * 2 + 1));
    if (want <= ctx->ds_cap) return PTX_OK;
    ulonglong2* nd = nullptr;
    CU(cudaMalloc((void**)&nd, want * sizeof(ulonglong2)));
    CU(cudaMemsetAsync(nd, 0, want * sizeof(ulonglong2), ctx->st));
    if (ctx->d_ds && (ctx->ds_records > 0 || ctx->ds_entries_bound > 0))
        launch_ds_rehash(ctx->d_ds, ctx->ds_cap, nd, 64 - log2_ceil(want), want - 1, ctx->ds_epoch, ctx->st);
    if (ctx->d_ds) {
        CU(cudaStreamSynchronize(ctx->st));
        cudaFree(ctx->d_ds);
    }
    ctx->d_ds = nd;
    ctx->ds_cap = want;
    return PTX_OK;
}
// A new pass over the reads: every slot of the id set becomes empty by moving on to the next epoch; the table is cleared for real
// only when the 8-bit epoch wraps (the memset of 512 MB per 10 M-record step was 4 % of it).
int ds_new_pass(ptx_ctx* ctx) {
    if (!ctx->d_ds) return PTX_OK;
    if (++ctx->ds_epoch > 255u) {
        ctx->ds_epoch = 1;
        CU(cudaMemsetAsync(ctx->d_ds, 0, ctx->ds_cap * sizeof(ulonglong2), ctx->st));
    }
    return PTX_OK;
}
int ds_ensure(ptx_ctx* ctx, int64_t more_records) { return ds_ensure_total(ctx, std::max(ctx->ds_records, ctx->ds_entries_bound) + more_records); }

// Which scatter variant k_apply<COVER> uses.  Measured on configs[1], configs[2], a 50 M-node graph and a 4,000-node "hot" graph
// (profiles/r2_scatter_bakeoff.md): the plain per-lane RED.ADD.64 wins on all four.  The 32 reads of a warp almost never share a
// node in the same step of their walks (match.any merges nothing: identical RED sector counts), a 2048-slot CTA table removes no
// global RED on a graph of a million nodes and its shared-memory atomics cost more than the REDs it saves on a graph of 4,000,
// and sort + segmented reduce moves 40x the DRAM bytes.  So: variant 0, unless PTX_SCATTER asks for another one (measurements).
int pick_scatter_variant(const ptx_ctx* ctx) {
    if (ctx->scatter_var >= 0 && ctx->scatter_var <= 3) return ctx->scatter_var;
    return 0;
}

IngestArgs make_args(ptx_ctx* ctx, const Chunk& ch) {
    IngestArgs a;
    memset(&a, 0, sizeof a);
    a.text = ch.buf + PRE;
    a.n_bytes = ch.n;
    a.padded_bytes = ch.padded;
    a.n_tiles = ch.n_tiles;
    a.tile_bytes = ch.tile_bytes;
    a.over_bytes = ch.over_bytes;
    a.no_sort = ctx->no_sort ? 1u : 0u;
    a.pf_dist = ctx->pf_waves > 0 ? (uint32_t)(ctx->pf_waves * ctx->n_sm * (ch.long_mode ? 3 : 7)) : 0u;  // resident CTAs per SM of k_ingest_l / k_ingest_s
    a.old_short = ctx->old_short ? 1u : 0u;
    a.long_new = ctx->long_new ? 1u : 0u;
    a.long_mode = ch.long_mode;
    a.micro_base = ch.tile_base;
    a.labels = ch.labels;
    a.labels_in = ctx->labels_in_n > 0 ? ctx->d_labels_in + ch.row_base : nullptr;
    a.meta_b = ch.meta_b;
    a.meta_a = ch.meta_a;
    a.hash_lo = ch.hash_lo;
    a.nodes = ch.nodes;
    a.nodes_cap = (uint32_t)std::min<int64_t>(ch.nodes_cap, 0xFFFFFFF0ll);
    a.cursors = reinterpret_cast<uint32_t*>(ch.cursors);

    a.box_ptr = ctx->comm ? ctx->d_box_ptr : nullptr;
    a.out_cursor = ctx->out_cursor;
    a.box_cap = ctx->box_cap;
    a.n_ranks = (uint32_t)ctx->n_ranks;
    a.rank = (uint32_t)ctx->rank;
    a.ranges = ranges_view(ctx);
    a.hist = ctx->d_hist;
    a.hist_copies = 1;
    a.hist_stride = 0;
    a.ds = ctx->d_ds;
    a.ds_shift = 64 - log2_ceil(ctx->ds_cap);
    a.ds_mask = ctx->ds_cap - 1;
    a.scatter_var = (uint32_t)pick_scatter_variant(ctx);
    a.pol_keep = (ctx->l2_hints & 1) ? 0x14F0000000000000ull : 0x1000000000000000ull;    // createpolicy evict_last / evict_normal, fraction 1.0
    a.pol_stream = (ctx->l2_hints & 2) ? 0x12F0000000000000ull : 0x1000000000000000ull;  // evict_first
    a.pol_ds = (ctx->l2_hints & 4) ? 0x12F0000000000000ull : 0x1000000000000000ull;
    a.ds_epoch = ctx->ds_epoch;
    a.pair_key = ctx->d_pair_key;
    a.pair_val = ctx->d_pair_val;
    a.flags = ctx->d_flags;
    a.err = ctx->d_err;
    const GraphDev& g = ctx->g;
    a.ninfo = g.ninfo;
    a.bases = g.bases;
    a.bits = g.bits;
    a.tt = g.T > 0 ? g.tt : nullptr;
    a.tt_mask = g.tt_mask;
    a.trio_bases = g.trio_bases;
    if (ch.single_pass) {
        a.micro_base = nullptr;
        a.tile_info = ch.tile_info;
        a.row_key = ch.row_key;
        a.slots_cap = (uint32_t)std::min<int64_t>(ch.slots_cap, 0xFFFFFFF0ll);
        a.hist = ch.chunk_hist;
        a.hist_copies = ch.hist_copies;
        a.hist_stride = (uint32_t)(std::max<size_t>(ctx->sp.size(), 1) * 4);
    }
    return a;
}

// bytes readable after `text`: whole 32 KB blocks covering n, one spare block (a tile of any size that
// starts inside the text ends before it), and the staging overhang
size_t padded_text_bytes(size_t n) {
    size_t blocks = (n + MAX_TILE - 1) / MAX_TILE;
    return (blocks + 1) * (size_t)MAX_TILE + OVER;
}

// a text buffer able to hold `cap` bytes: [PRE '\n'][text][padding]; from the pool of released ones if one fits
int text_alloc(ptx_ctx* ctx, Chunk& ch, size_t cap) {
    int best = -1;
    for (size_t i = 0; i < ctx->text_pool.size(); ++i) {
        const auto& t = ctx->text_pool[i];
        if (t.cap >= cap && t.cap <= 2 * cap + (1u << 20) && (best < 0 || t.cap < ctx->text_pool[best].cap)) best = (int)i;
    }
    if (best >= 0) {
        ch.buf = ctx->text_pool[best].buf;
        ch.cap = ctx->text_pool[best].cap;
        ch.copied = ctx->text_pool[best].copied;
        ctx->text_pool.erase(ctx->text_pool.begin() + best);
        return PTX_OK;
    }
    ch.cap = cap;
    const size_t total = PRE + padded_text_bytes(cap);
    const cudaError_t e = cudaMalloc((void**)&ch.buf, total);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ch.buf = nullptr;
        return fail(ctx, PTX_E_NOMEM, "out of device memory for a %zu MB GAF text buffer: the record tables of the input ingested so far stay resident "
                                      "(about 0.6 bytes per GAF byte) - split the input over more GPUs or contexts", total >> 20);
    }
    // the newline padding in front of the text is written once, on the COPY stream: the text copies are queued behind it
    // there, and the kernels wait for `copied` - no host synchronisation between a call's copies and the previous call's kernels
    CU(cudaMemsetAsync(ch.buf, '\n', PRE, ctx->copy_st));
    CU(cudaEventCreateWithFlags(&ch.copied, cudaEventDisableTiming));
    return PTX_OK;
}

// a chunk (record tables + text buffer) for `cap` text bytes
int chunk_alloc(ptx_ctx* ctx, Chunk& ch, size_t cap) {
    // reuse a released chunk of similar size (ptx_reset keeps them): no cudaMalloc on the steady-state path
    int best = -1;
    for (size_t i = 0; i < ctx->pool.size(); ++i) {
        const Chunk& c = ctx->pool[i];
        const bool fits = c.buf == nullptr ? true : (c.cap >= cap && c.cap <= 2 * cap + (1u << 20));
        if (fits && (best < 0 || (c.buf != nullptr && ctx->pool[best].buf == nullptr))) best = (int)i;
    }
    if (best >= 0) {
        ch = ctx->pool[best];
        ctx->pool.erase(ctx->pool.begin() + best);
        ch.n = 0; ch.n_records = 0; ch.ingested = false; ch.covered = false; ch.pending = false; ch.host_text = false;
    } else {
        ch = Chunk();
    }
    if (ch.buf == nullptr) return text_alloc(ctx, ch, cap);
    return PTX_OK;
}

// hand the text buffer of a resolved chunk back (see ptx_ctx::text_pool)
void chunk_release_text(ptx_ctx* ctx, Chunk& ch) {
    if (!ch.buf) return;
    ctx->text_pool.push_back({ch.buf, ch.cap, ch.copied});
    ch.buf = nullptr;
    ch.copied = nullptr;
    ch.cap = 0;
}

// bump allocation from the grow-only finalize scratch arena (256-byte aligned); reset with scratch_off = 0
int scratch_reserve(ptx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_cap) return PTX_OK;
    CU(cudaStreamSynchronize(ctx->st));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_cap = 0;
    CU(cudaMalloc((void**)&ctx->scratch, bytes));
    ctx->scratch_cap = bytes;
    return PTX_OK;
}
template <class T>
T* scratch_take(ptx_ctx* ctx, size_t n) {
    size_t off = (ctx->scratch_off + 255) & ~(size_t)255;
    ctx->scratch_off = off + n * sizeof(T);
    return reinterpret_cast<T*>(ctx->scratch + off);
}

void chunk_free(Chunk& ch) {
    dfree(ch.buf);
    dfree(ch.tile_count);
    dfree(ch.tile_base);
    dfree(ch.scan_scratch);
    dfree(ch.labels);
    dfree(ch.meta_b); dfree(ch.meta_a); dfree(ch.hash_lo); dfree(ch.nodes); dfree(ch.cursors);
    dfree(ch.tile_info); dfree(ch.row_key); dfree(ch.chunk_hist);
    if (ch.h_cur) cudaFreeHost(ch.h_cur);
    ch.h_cur = nullptr;
    if (ch.done) cudaEventDestroy(ch.done);
    ch.done = nullptr;
    if (ch.copied) cudaEventDestroy(ch.copied);
    ch.copied = nullptr;
}

int xchg_ensure(ptx_ctx* ctx, int64_t records);
int first_rows_exchange(ptx_ctx* ctx, const std::vector<unsigned long long>& counts);

// (re)allocate the record table of a chunk for `slots` line slots
int chunk_table_ensure(ptx_ctx* ctx, Chunk& ch, int64_t slots) {
    if (ch.slots_cap >= slots) return PTX_OK;
    ++ctx->n_table_allocs;
    dfree(ch.meta_b); dfree(ch.meta_a); dfree(ch.hash_lo); dfree(ch.row_key);
    // headroom: single-pass estimates move a little from chunk to chunk; a cudaFree/cudaMalloc in the middle of a
    // stream of chunks would synchronise the device (measured: e2e 22.6 -> 27-70 ms per step)
    const size_t cap = (size_t)std::max<int64_t>(slots + slots / 8, 1);
    CU(cudaMalloc((void**)&ch.meta_b, cap * sizeof(uint4)));
    CU(cudaMalloc((void**)&ch.meta_a, cap * sizeof(longlong2)));
    CU(cudaMalloc((void**)&ch.hash_lo, cap * sizeof(unsigned long long)));
    CU(cudaMalloc((void**)&ch.row_key, cap * sizeof(uint16_t)));
    ch.slots_cap = (int64_t)cap;
    return PTX_OK;
}

void chunk_pick_tile(ptx_ctx* ctx, Chunk& ch, double mean_line, bool exact) {
    // long lines (HiFi/ONT walks): even the largest tile holds fewer lines than threads -> one warp per record
    ch.long_mode = mean_line >= LONG_LINE_BYTES ? 1u : 0u;
    if (ctx->force_long >= 0) ch.long_mode = (uint32_t)ctx->force_long;
    // tile size: about one record per thread.  k_ingest_s takes any multiple of 512 bytes; the count pass numbers the rows
    // per 4 KB micro-tile, so a chunk that ran it (and the two older kernels) gets whole micro-tiles
    const bool short_new = !ch.long_mode && !ctx->old_short;
    const bool fine = !exact && short_new;
    const uint32_t threads = short_new ? SHORT_THREADS : INGEST_THREADS;
    const uint32_t gran = fine ? 128u : MICRO;
    uint32_t tile = (uint32_t)(0.97 * threads * mean_line) / gran * gran;
    tile = std::min<uint32_t>(MAX_TILE, std::max<uint32_t>(MICRO, tile));
    if (ch.long_mode && ctx->long_new) tile = std::min<uint32_t>(tile, ctx->long_tile_max);
    if (ctx->force_rows > 0) tile = (uint32_t)std::min(8, ctx->force_rows) * MICRO;
    if (ctx->force_tile > 0 && fine) tile = std::min<uint32_t>(MAX_TILE, std::max<uint32_t>(MICRO, (uint32_t)ctx->force_tile / 128u * 128u));
    // the window behind the tile holds the tail of its last line: four mean lines, at least 512 bytes (longer lines are re-read from global memory)
    ch.over_bytes = std::min<uint32_t>(OVER, std::max<uint32_t>(512u, ((uint32_t)(4.0 * mean_line) + 511u) / 512u * 512u));
    ch.tile_bytes = tile;
    ch.n_tiles = (uint32_t)((ch.n + tile - 1) / tile);
}

int chunk_common_begin(ptx_ctx* ctx, Chunk& ch) {
    // pad behind the text with newlines (terminates an unterminated last line; padding holds no records)
    ch.padded = padded_text_bytes(ch.n);
    CU(cudaMemsetAsync(ch.buf + PRE + ch.n, '\n', ch.padded - ch.n, ctx->st));
    ch.n_micro = (uint32_t)((ch.padded - OVER) / MICRO);
    if (ch.n >= (1ull << 32)) return fail(ctx, PTX_E_INVALID, "a chunk must be smaller than 4 GiB (split the input)");
    if (!ch.cursors) {
        CU(cudaMalloc((void**)&ch.cursors, 8 * sizeof(unsigned long long)));
        CU(cudaHostAlloc((void**)&ch.h_cur, 8 * sizeof(uint32_t), cudaHostAllocDefault));
        CU(cudaEventCreateWithFlags(&ch.done, cudaEventDisableTiming));
    }
    CU(cudaMemsetAsync(ch.cursors, 0, 8 * sizeof(unsigned long long), ctx->st));
    return PTX_OK;
}

// CSR node slots of a chunk: a walk node takes at least two bytes of text; the long-line kernel reserves
// (bytes from column 6 to the end of the line) / 2 + 1 slots per record instead of counting the nodes first
int chunk_nodes_ensure(ptx_ctx* ctx, Chunk& ch, bool exact) {
    const int64_t bound = (int64_t)(ch.n / 2 + 16) + (ch.long_mode ? ch.slots_cap : 0);  // always enough
    int64_t need = bound;
    // single-pass chunks of k_ingest_s: from the walk nodes per byte of the chunks seen so far, with a quarter of slack; the kernel
    // abandons the chunk (redone exactly) instead of writing beyond the buffer
    if (!exact && !ch.long_mode && !ctx->old_short && ctx->seen_nodes_per_byte > 0)
        need = std::min<int64_t>(bound, (int64_t)((double)ch.n * ctx->seen_nodes_per_byte * 1.25) + 65536);
    if (ch.nodes_cap < need) {
        dfree(ch.nodes);
        ch.nodes_cap = need + need / 8;
        const cudaError_t e = cudaMalloc((void**)&ch.nodes, (size_t)ch.nodes_cap * sizeof(uint32_t));
        if (e != cudaSuccess) {
            cudaGetLastError();
            ch.nodes = nullptr;
            ch.nodes_cap = 0;
            return fail(ctx, PTX_E_NOMEM, "out of device memory for the walk table of a chunk (%lld MB): the record tables of the input ingested so far stay "
                                          "resident (about 0.6 bytes per GAF byte) - split the input over more GPUs or contexts", (long long)(need >> 18));
        }
    }
    return PTX_OK;
}

// Exact pass: a count kernel numbers the GAF rows first (records per 4 KB + scan), the host reads the totals back
// and sizes everything exactly.  Used for the first chunk of a ctx, with ptx_ingest_labels, and to redo a chunk
// whose single-pass estimate was too small.
int chunk_process_exact(ptx_ctx* ctx, Chunk& ch) {
    int rc = chunk_common_begin(ctx, ch);
    if (rc) return rc;
    if (ch.tiles_cap < ch.n_micro) {
        dfree(ch.tile_count);
        dfree(ch.tile_base);
        dfree(ch.scan_scratch);
        CU(cudaMalloc((void**)&ch.tile_count, ch.n_micro * sizeof(uint32_t)));
        CU(cudaMalloc((void**)&ch.tile_base, ((size_t)ch.n_micro + 1) * sizeof(uint64_t)));
        CU(cudaMalloc((void**)&ch.scan_scratch, ((size_t)ch.n_micro / 2048 + 4) * sizeof(uint64_t)));
        ch.tiles_cap = ch.n_micro;
    }
    ev_begin(ctx, ctx->ev_count);
    launch_count_records(ch.buf + PRE, ch.n, ch.n_micro, ch.tile_count, ch.cursors + 4, ctx->st);
    launch_scan_u32(ch.tile_count, ch.tile_base, ch.n_micro, ch.scan_scratch, ctx->st);
    ev_end(ctx, ctx->ev_count);
    uint64_t total = 0, slots = 0;
    CU(cudaMemcpyAsync(&total, ch.tile_base + ch.n_micro, sizeof total, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&slots, ch.cursors + 4, sizeof slots, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ch.n_records = (int64_t)total;
    ch.n_slots = (int64_t)slots;
    if (slots > 0xFFFFFFF0ull) return fail(ctx, PTX_E_INVALID, "more than 2^32 lines in one chunk");
    if ((rc = chunk_table_ensure(ctx, ch, ch.n_slots))) return rc;
    const double mean_line = (double)ch.n / (double)std::max<uint64_t>(total, 1);
    chunk_pick_tile(ctx, ch, mean_line, true);
    if ((rc = chunk_nodes_ensure(ctx, ch, true))) return rc;
    if (!ch.labels || ch.labels_cap < (int64_t)total) {
        dfree(ch.labels);
        CU(cudaMalloc((void**)&ch.labels, std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
        ch.labels_cap = (int64_t)std::max<uint64_t>(total, 1);
    }
    ch.row_base = ctx->total_records;
    if (ctx->labels_in_n > 0 && ch.row_base + ch.n_records > ctx->labels_in_n)
        return fail(ctx, PTX_E_STATE, "ptx_ingest_labels supplied fewer labels than the GAF has rows");
    if ((rc = ds_ensure(ctx, ch.n_records + ctx->pending_records))) return rc;
    ctx->ds_records += ch.n_records;
    if (ctx->comm) {
        rc = xchg_ensure(ctx, std::max<int64_t>(ctx->ds_records + ctx->pending_records, ctx->reserve_records));
        if (rc) return rc;
    }
    ch.single_pass = false;
    ch.labels_ready = true;
    IngestArgs a = make_args(ctx, ch);
    // multi-GPU: the coverage pass is deferred to ptx_finalize, where it overlaps the id-group exchange
    const bool cover = ctx->graphs_committed && ctx->g.N > 0 && !ctx->comm && a.scatter_var != 3;  // variant 3 sorts per chunk in ptx_finalize
    ev_begin(ctx, ctx->ev_ingest);
    launch_ingest(a, ctx->st);
    ev_end(ctx, ctx->ev_ingest);
    ev_begin(ctx, ctx->ev_apply);
    launch_apply(a, (uint32_t)ch.n_slots, MODE_CLASSIFY | (cover ? MODE_COVER : 0), ctx->st);
    ev_end(ctx, ctx->ev_apply);
    CU(cudaGetLastError());
    ch.ingested = true;
    ch.covered = cover;
    ch.pending = false;
    ctx->total_records += ch.n_records;
    ctx->dirty = true;
    if (total > 0) { ctx->seen_mean_line = mean_line; ctx->seen_slots_per_row = (double)slots / (double)total; }
    return PTX_OK;
}

// Single pass: the text is read ONCE and the host does not wait.  Tile size and table capacity come from the line
// statistics of the chunks seen so far; k_ingest numbers the rows per tile (tile_info + row_key) instead of needing
// a global row prefix, k_apply takes its entry count from the device cursor.  The counts come back asynchronously
// and are looked at by chunk_resolve (finalize / getters); a too small estimate abandons the chunk on the device
// (nothing of it is counted) and chunk_resolve redoes it with chunk_process_exact.
int chunk_process_single(ptx_ctx* ctx, Chunk& ch) {
    int rc = chunk_common_begin(ctx, ch);
    if (rc) return rc;
    chunk_pick_tile(ctx, ch, ctx->seen_mean_line, false);
    const double est_rows = (double)ch.n / ctx->seen_mean_line;
    const int64_t est_slots = (int64_t)(est_rows * ctx->seen_slots_per_row * 1.25) + 4096;
    if (est_slots > 0xFFFFFFF0ll) return chunk_process_exact(ctx, ch);
    if ((rc = chunk_table_ensure(ctx, ch, est_slots))) return rc;
    if ((rc = chunk_nodes_ensure(ctx, ch, false))) return rc;
    if (ch.tile_info_cap < ch.n_tiles) {
        dfree(ch.tile_info);
        CU(cudaMalloc((void**)&ch.tile_info, (size_t)ch.n_tiles * sizeof(uint4)));
        ch.tile_info_cap = ch.n_tiles;
    }
    const size_t S4 = std::max<size_t>(ctx->sp.size(), 1) * 4;
    // the species counts of a chunk go to HIST_COPIES private copies (tile t adds to copy t % copies; k_hist_merge sums them): an abundant
    // species is one hot address per copy instead of one for the whole grid
    ch.hist_copies = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, ((size_t)8 << 20) / (S4 * sizeof(unsigned long long))));
    if (ch.hist_cap < S4 * ch.hist_copies) {
        dfree(ch.chunk_hist);
        CU(cudaMalloc((void**)&ch.chunk_hist, S4 * ch.hist_copies * sizeof(unsigned long long)));
        ch.hist_cap = S4 * ch.hist_copies;
    }
    CU(cudaMemsetAsync(ch.chunk_hist, 0, S4 * ch.hist_copies * sizeof(unsigned long long), ctx->st));
    ch.est_records = (int64_t)(est_rows * 1.25) + 4096;
    if ((rc = ds_ensure(ctx, ctx->pending_records + ch.est_records))) return rc;
    if (ctx->comm) {
        rc = xchg_ensure(ctx, std::max<int64_t>(ctx->ds_records + ctx->pending_records + ch.est_records, ctx->reserve_records));
        if (rc) return rc;
    }
    ctx->pending_records += ch.est_records;
    ch.single_pass = true;
    ch.labels_ready = false;
    ch.n_slots = ch.slots_cap;  // until resolved
    IngestArgs a = make_args(ctx, ch);
    const bool cover = ctx->graphs_committed && ctx->g.N > 0 && !ctx->comm && a.scatter_var != 3;
    ev_begin(ctx, ctx->ev_ingest);
    launch_ingest(a, ctx->st);
    ev_end(ctx, ctx->ev_ingest);
    ev_begin(ctx, ctx->ev_apply);
    launch_apply(a, ENTRIES_FROM_DEVICE, MODE_CLASSIFY | (cover ? MODE_COVER : 0), ctx->st);
    ev_end(ctx, ctx->ev_apply);
    launch_hist_merge(ch.chunk_hist, ctx->d_hist, (uint32_t)S4, ch.hist_copies, a.cursors, ctx->st);
    CU(cudaMemcpyAsync(ch.h_cur, a.cursors, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaEventRecord(ch.done, ctx->st));
    CU(cudaGetLastError());
    ch.ingested = true;
    ch.covered = cover;
    ch.pending = true;
    ctx->dirty = true;
    return PTX_OK;
}

// Look at the counts of the single-pass chunks; redo abandoned ones exactly.  blocking: wait for every chunk
// (finalize, getters); otherwise only take what has already arrived (keeps the line statistics fresh while streaming).
int chunks_resolve(ptx_ctx* ctx, bool blocking = true) {
    for (size_t ci = 0; ci < ctx->chunks.size(); ++ci) {
        Chunk& ch = ctx->chunks[ci];
        if (!ch.pending) continue;
        if (blocking) CU(cudaEventSynchronize(ch.done));
        else if (cudaEventQuery(ch.done) != cudaSuccess) { cudaGetLastError(); break; }  // chunks resolve in stream (= GAF) order: rows are numbered from it
        ch.pending = false;
        ctx->pending_records -= ch.est_records;
        if (ch.h_cur[3]) {  // estimate too small (or a tile with more than REC_CAP lines): nothing was counted
            ch.ingested = false;
            ch.covered = false;
            int rc = chunk_process_exact(ctx, ch);
            if (rc) return rc;
            continue;
        }
        ch.n_slots = (int64_t)ch.h_cur[0];
        ch.n_records = (int64_t)ch.h_cur[2];
        ch.row_base = ctx->total_records;
        ctx->total_records += ch.n_records;
        ctx->ds_records += ch.n_records;
        if (ch.n_records > 0) {
            ctx->seen_mean_line = (double)ch.n / (double)ch.n_records;
            ctx->seen_slots_per_row = (double)ch.n_slots / (double)ch.n_records;
            ctx->seen_nodes_per_byte = (double)ch.h_cur[1] / (double)ch.n;
        }
        // the chunk is complete and will never be parsed again: unless ptx_equal_length may still want its lines (the first 1000
        // non-U rows, profile.rs:311-322), its text buffer goes back to the pool for the pieces still to come
        if (ch.host_text && ctx->labelled_rows_known >= 1000 && !ctx->keep_text) chunk_release_text(ctx, ch);
        ctx->labelled_rows_known += (int64_t)ch.h_cur[4];
    }
    return PTX_OK;
}

// labels[] of a single-pass chunk in GAF row order, built from the record table on demand
int chunk_labels_materialize(ptx_ctx* ctx, Chunk& ch) {
    if (ch.labels_ready || ch.n_records == 0) return PTX_OK;
    if (!ch.labels || ch.labels_cap < ch.n_records) {
        dfree(ch.labels);
        CU(cudaMalloc((void**)&ch.labels, (size_t)std::max<int64_t>(ch.n_records, 1) * sizeof(uint32_t)));
        ch.labels_cap = std::max<int64_t>(ch.n_records, 1);
    }
    uint32_t* rows = nullptr;
    uint64_t *off = nullptr, *scr = nullptr;
    int rc;
    if ((rc = dalloc(ctx, &rows, (size_t)ch.n_tiles, false)) || (rc = dalloc(ctx, &off, (size_t)ch.n_tiles + 1, false)) ||
        (rc = dalloc(ctx, &scr, (size_t)ch.n_tiles / 2048 + 4, false)))
        return rc;
    launch_tile_rows(ch.tile_info, rows, ch.n_tiles, ctx->st);
    launch_scan_u32(rows, off, ch.n_tiles, scr, ctx->st);
    launch_labels_from_table(ch.tile_info, off, ch.meta_b, ch.row_key, ch.labels, ch.n_tiles, ctx->st);
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(rows); cudaFree(off); cudaFree(scr);
    ch.labels_ready = true;
    return PTX_OK;
}

// classify (+ optimistic coverage) pass over one chunk whose text is resident
int chunk_process(ptx_ctx* ctx, Chunk& ch) {
    if (ctx->cov_reduced) return fail(ctx, PTX_E_STATE, "multi-GPU: coverage already reduced by ptx_finalize; ptx_reset before ingesting more");
    if (ch.n == 0) { ch.n_tiles = 0; ch.ingested = true; ch.covered = true; ch.pending = false; ch.n_records = 0; ch.n_slots = 0; return PTX_OK; }
    {
        int rc = chunks_resolve(ctx, false);
        if (rc) return rc;
    }
    const bool single = ctx->single_pass_ok && ctx->seen_mean_line > 0 && ctx->labels_in_n == 0;
    return single ? chunk_process_single(ctx, ch) : chunk_process_exact(ctx, ch);
}

int zero_coverage(ptx_ctx* ctx) {
    GraphDev& g = ctx->g;
    if (g.N <= 0) return PTX_OK;
    CU(cudaMemsetAsync(g.bases, 0, g.N * sizeof(unsigned long long), ctx->st));
    launch_ninfo_full(g.ninfo, g.full, g.N, 1, ctx->st);
    CU(cudaMemsetAsync(g.bits, 0, g.n_bit_words * sizeof(uint32_t), ctx->st));
    if (g.T > 0) CU(cudaMemsetAsync(g.trio_bases, 0, g.T * sizeof(unsigned long long), ctx->st));
    CU(cudaMemsetAsync(ctx->d_err, 0, std::max<size_t>(ctx->sp.size(), 1) * sizeof(uint32_t), ctx->st));
    return PTX_OK;
}

void free_graph(ptx_ctx* ctx) {
    GraphDev& g = ctx->g;
    dfree(g.len); dfree(g.bit_off); dfree(g.bases); dfree(g.full); dfree(g.ninfo); dfree(g.bits); dfree(g.cov);
    dfree(g.pnode); dfree(g.poff); dfree(g.path_len_sum); dfree(g.path_cov_sum);
    dfree(g.trio_key); dfree(g.trio_len); dfree(g.trio_owner); dfree(g.trio_bases); dfree(g.trio_start);
    dfree(g.hap_nz); dfree(g.tt);
    g = GraphDev();
}

int check_species(ptx_ctx* ctx, int s, bool need_graph) {
    if (!ctx) return PTX_E_INVALID;
    if (s < 0 || s >= (int)ctx->sp.size()) return fail(ctx, PTX_E_RANGE, "species index %d out of range", s);
    if (need_graph && !ctx->sp[s].has_graph) return fail(ctx, PTX_E_NO_GRAPH, "species %d has no committed graph", s);
    return PTX_OK;
}

int check_results(ptx_ctx* ctx, int s) {
    int rc = check_species(ctx, s, true);
    if (rc) return rc;
    if (ctx->dirty) return fail(ctx, PTX_E_STATE, "call ptx_finalize before reading results");
    if (s < (int)ctx->h_err.size() && (ctx->h_err[s] & 1u))
        return fail(ctx, PTX_E_START_GT_LEN, "species %s: read start is bigger than node len (profile.rs:854)",
                    ctx->sp[s].taxid.c_str());
    return PTX_OK;
}

int nccl_check(ptx_ctx* ctx, int r, const char* what) {
    if (r == 0) return PTX_OK;
    return fail(ctx, PTX_E_NCCL, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}


// multi-GPU: make the per-owner boxes large enough for `records` labelled records on this rank (uniform hash:
// records/P each, 25 % slack; k_apply flags an overflow and ptx_finalize then falls back to the table scan)
int xchg_ensure(ptx_ctx* ctx, int64_t records) {
    const uint64_t P = (uint64_t)ctx->n_ranks;
    uint64_t need = (uint64_t)records / P + (uint64_t)records / (4 * P) + 4096;
    if (ctx->test_box_cap > 0 && ctx->box_cap == 0) { need = (uint64_t)ctx->test_box_cap; ctx->test_box_cap = 0; }
    if (!ctx->xs) {
        CU(cudaStreamCreateWithFlags(&ctx->xs, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_x0, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_x1, cudaEventDisableTiming));
        // [0,P) box cursors, [P] "an entry was dropped", then the all-gathered (P+1) x P matrix and 2P merge parameters
        int rc = dalloc(ctx, &ctx->out_cursor, (size_t)(P + 1) * (P + 1) + 2 * P + 8);
        if (rc) return rc;
        if (!ctx->d_box_ptr && (rc = dalloc(ctx, &ctx->d_box_ptr, (size_t)P))) return rc;
        ctx->box_sent.assign(P, 0);
        ctx->recv_done.assign(P, 0);
    }
    if (ctx->p2p || need <= ctx->box_cap) return PTX_OK;  // peer-memory boxes have a fixed size (overflow -> restart without them)
    const uint64_t cap = need + need / 8;
    ulonglong2* nout = nullptr;
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaStreamSynchronize(ctx->xs));
    CU(cudaMalloc((void**)&nout, P * cap * sizeof(ulonglong2)));
    if (ctx->outbox)  // keep what earlier chunks already wrote
        for (uint64_t r = 0; r < P; ++r)
            CU(cudaMemcpyAsync(nout + r * cap, ctx->outbox + r * ctx->box_cap, ctx->box_cap * sizeof(ulonglong2), cudaMemcpyDeviceToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    dfree(ctx->outbox);
    ctx->outbox = nout;
    ctx->box_cap = cap;
    std::vector<ulonglong2*> ptrs(P);
    for (uint64_t r = 0; r < P; ++r) ptrs[r] = nout + r * cap;
    CU(cudaMemcpy(ctx->d_box_ptr, ptrs.data(), P * sizeof(ulonglong2*), cudaMemcpyHostToDevice));
    return PTX_OK;
}

// ids whose merged state on their owner rank is DS_MIXED go back to every rank (rare path; synchronising).  A rank
// keeps only the ids it owns, so the returned ids are inserted (as MIXED) where they are not present.
int return_mixed_ids(ptx_ctx* ctx, cudaStream_t st) {
    const int P = ctx->n_ranks;
    int rc;
    unsigned long long* d_tmp = nullptr;
    if ((rc = dalloc(ctx, &d_tmp, (size_t)P + 2))) return rc;
    unsigned long long* d_nmix = d_tmp;
    unsigned long long* d_allmix = d_tmp + 1;
    launch_ds_collect_mixed(ctx->d_ds, ctx->ds_cap, ctx->ds_epoch, d_nmix, nullptr, 0, st);
    if ((rc = nccl_check(ctx, g_nccl.AllGather(d_nmix, d_allmix, 1, ncclUint64, ctx->comm, st), "ncclAllGather(mixed counts)"))) return rc;
    std::vector<unsigned long long> nmix(P);
    CU(cudaMemcpyAsync(nmix.data(), d_allmix, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsigned long long mx = 0, total = 0;
    for (auto v : nmix) { mx = std::max(mx, v); total += v; }
    if (mx > 0) {
        ulonglong2* mine = nullptr;
        ulonglong2* everyone = nullptr;
        if ((rc = dalloc(ctx, &mine, (size_t)mx)) || (rc = dalloc(ctx, &everyone, (size_t)mx * P))) return rc;
        CU(cudaMemsetAsync(d_nmix, 0, sizeof(unsigned long long), st));
        launch_ds_collect_mixed(ctx->d_ds, ctx->ds_cap, ctx->ds_epoch, d_nmix, mine, mx, st);
        if ((rc = nccl_check(ctx, g_nccl.AllGather(mine, everyone, mx * 2, ncclUint64, ctx->comm, st), "ncclAllGather(mixed ids)"))) return rc;
        ctx->ds_entries_bound += (int64_t)total;
        if ((rc = ds_ensure_total(ctx, ctx->ds_entries_bound))) return rc;
        launch_ds_apply_mixed(everyone, mx * P, ctx->d_ds, 64 - log2_ceil(ctx->ds_cap), ctx->ds_cap - 1, ctx->ds_epoch, ctx->d_flags, st);
        CU(cudaStreamSynchronize(st));
        cudaFree(mine);
        cudaFree(everyone);
    }
    CU(cudaStreamSynchronize(st));
    cudaFree(d_tmp);
    return PTX_OK;
}

// Id groups across ranks (profile.rs:369-378 / 406-437 are keyed by read id, which ignores shard boundaries).
// k_apply kept the ids this rank owns in its id set and appended every other record's {hash, state} to the outbox
// of the owner.  Here: the box fills are all-gathered (one small collective + readback, the only host
// synchronisation), then - on the side stream `xs`, while the main stream runs the coverage pass - the new part
// of every box travels with grouped ncclSend/ncclRecv, the owner merges what it received into its id set and
// the repeat/mixed flags are max-reduced.  ptx_finalize joins on ev_x1.
int exchange_begin(ptx_ctx* ctx, const std::function<int()>& overlap) {
    const int P = ctx->n_ranks;
    int rc;
    unsigned long long* d_cur = ctx->out_cursor;           // [P] cursors + [P] overflow marker (set below)
    unsigned long long* d_all = ctx->out_cursor + (P + 1);  // [(P+1) * P]
    unsigned long long* d_par = d_all + (size_t)(P + 1) * P;  // [2P] merge offsets, counts
    std::vector<unsigned long long> all((size_t)(P + 1) * P);
    for (int attempt = 0;; ++attempt) {
        // min(non-U rows of this rank, 1000) rides in the upper bits of the overflow marker (ptx_equal_length needs the first 1000 of the whole GAF)
        launch_count_labelled(ctx->d_hist, (uint32_t)ctx->sp.size(), d_cur + P, ctx->st);
        if ((rc = nccl_check(ctx, g_nccl.AllGather(d_cur, d_all, P + 1, ncclUint64, ctx->comm, ctx->st), "ncclAllGather(box fills)"))) return rc;
        CU(cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaEventRecord(ctx->ev_x0, ctx->st));
        // the coverage pass does not depend on the exchange: it is queued before the host waits for the fills, so the GPU goes from
        // the all-gather straight into it while the host reads the fills and issues the merge on the side stream
        if (attempt == 0 && overlap && (rc = overlap())) return rc;
        CU(cudaEventSynchronize(ctx->ev_x0));
        // all[q*(P+1) + r] = entries rank q has appended for rank r so far, all[q*(P+1) + P] != 0 if a box of rank q was
        // full and entries were dropped.  Every rank sees the same matrix, so all of them take the same branch.
        unsigned long long worst = 0;
        bool over = false;
        for (int q = 0; q < P; ++q) {
            if (all[(size_t)q * (P + 1) + P] & 1ull) over = true;  // rank q dropped an entry
            for (int r = 0; r < P; ++r) worst = std::max(worst, all[(size_t)q * (P + 1) + r]);
        }
        if (!over) break;
        if (attempt >= 2) return fail(ctx, PTX_E_STATE, "id-group exchange: outboxes still overflow after being enlarged");
        // rare: some rank met far more foreign ids than planned.  Start the id sets over everywhere with boxes that
        // hold the fullest one seen, and rebuild them from the record tables (no text is re-read).
        if (ctx->p2p) {  // the peer-memory slices cannot grow: continue with local outboxes + ncclSend/ncclRecv
            ctx->p2p = false;
            dfree(ctx->outbox);
            ctx->box_cap = 0;
        }
        if ((rc = xchg_ensure(ctx, (int64_t)((worst + worst / 4) * (unsigned long long)P)))) return rc;
        if ((rc = ds_new_pass(ctx))) return rc;
        CU(cudaMemsetAsync(ctx->d_flags, 0, 3 * sizeof(uint32_t), ctx->st));
        CU(cudaMemsetAsync(d_cur, 0, (P + 1) * sizeof(unsigned long long), ctx->st));
        std::fill(ctx->box_sent.begin(), ctx->box_sent.end(), 0ull);
        std::fill(ctx->recv_done.begin(), ctx->recv_done.end(), 0ull);
        for (auto& ch : ctx->chunks) {
            if (!ch.ingested || ch.n_tiles == 0) continue;
            IngestArgs a = make_args(ctx, ch);
            launch_apply(a, (uint32_t)ch.n_slots, MODE_CLASSIFY, ctx->st);
        }
    }
    if (P > 1 && !ctx->first_rows_global) {  // rank 0's batch holds fewer than 1000 non-U rows (tiny inputs): the first 1000 span ranks
        std::vector<unsigned long long> counts((size_t)P);
        for (int q = 0; q < P; ++q) counts[(size_t)q] = all[(size_t)q * (P + 1) + P] >> 8;
        if (counts[0] < 1000 && (rc = first_rows_exchange(ctx, counts))) return rc;
    }
    // what is new since the previous exchange (ptx_finalize may run more than once)
    std::vector<unsigned long long> send_n(P, 0), recv_n(P, 0), roff(P, 0), par(2 * (size_t)P, 0);
    unsigned long long n_recv = 0, max_recv = 0, recv_total = 0;
    for (int q = 0; q < P; ++q) {
        if (q == ctx->rank) continue;
        send_n[q] = all[(size_t)ctx->rank * (P + 1) + q] - ctx->box_sent[q];
        const unsigned long long cum = all[(size_t)q * (P + 1) + ctx->rank];
        recv_n[q] = cum - ctx->recv_done[q];
        roff[q] = ctx->p2p ? (unsigned long long)q * ctx->p2p_cap + ctx->recv_done[q] : n_recv;  // entries rank q stored in its slice of my inbox
        n_recv += recv_n[q];
        max_recv = std::max(max_recv, recv_n[q]);
        recv_total += cum;
    }
    // the id set now also holds the foreign ids this rank owns
    ctx->ds_entries_bound = std::max<int64_t>(ctx->ds_entries_bound, ctx->ds_records + (int64_t)recv_total);
    if ((rc = ds_ensure_total(ctx, ctx->ds_entries_bound))) return rc;
    if (!ctx->p2p && n_recv > ctx->inbox_cap) {
        CU(cudaStreamSynchronize(ctx->xs));
        dfree(ctx->inbox);
        ctx->inbox_cap = n_recv + n_recv / 8 + 1024;
        CU(cudaMalloc((void**)&ctx->inbox, ctx->inbox_cap * sizeof(ulonglong2)));
    }
    int nb = 0;
    for (int q = 0; q < P; ++q) {
        if (q == ctx->rank || recv_n[q] == 0) continue;
        par[nb] = roff[q];
        par[(size_t)P + nb] = recv_n[q];
        ++nb;
    }
    // the side stream continues behind the fills all-gather (= behind this rank's k_apply launches and, through the collective, behind
    // the other ranks' stores into this rank's inbox), not behind the coverage pass queued after it
    CU(cudaStreamWaitEvent(ctx->xs, ctx->ev_x0, 0));
    cudaStream_t xs = ctx->xs;
    CU(cudaMemcpyAsync(d_par, par.data(), par.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, xs));
    Trace trx(xs);
    if (!ctx->p2p) {
        g_nccl.GroupStart();
        for (int q = 0; q < P; ++q) {
            if (q == ctx->rank) continue;
            if (send_n[q]) g_nccl.Send(ctx->outbox + (uint64_t)q * ctx->box_cap + ctx->box_sent[q], send_n[q] * 2, ncclUint64, q, ctx->comm_x, xs);
            if (recv_n[q]) g_nccl.Recv(ctx->inbox + roff[q], recv_n[q] * 2, ncclUint64, q, ctx->comm_x, xs);
        }
        if ((rc = nccl_check(ctx, g_nccl.GroupEnd(), "ncclSend/Recv(id boxes)"))) return rc;
        trx.mark("xchg send/recv");
    }
    // peer-memory boxes: the entries are already here - k_apply of the other ranks stored them into this rank's inbox
    // over NVLink; the all-gather above ordered this point behind those kernels
    launch_ds_merge_boxes(ctx->p2p ? ctx->p2p_inbox : ctx->inbox, d_par, d_par + P, (uint32_t)nb, max_recv, ctx->d_ds, 64 - log2_ceil(ctx->ds_cap), ctx->ds_cap - 1, ctx->ds_epoch, ctx->d_flags, xs);
    trx.mark("xchg merge");
    if ((rc = nccl_check(ctx, g_nccl.AllReduce(ctx->d_flags, ctx->d_flags, 2, ncclUint32, ncclMax, ctx->comm_x, xs), "ncclAllReduce(flags)"))) return rc;
    trx.mark("xchg flags");
    CU(cudaEventRecord(ctx->ev_x1, xs));
    for (int q = 0; q < P; ++q) {
        if (q == ctx->rank) continue;
        ctx->box_sent[q] += send_n[q];
        ctx->recv_done[q] += recv_n[q];
    }
    return PTX_OK;
}

// Peer-memory id boxes (all ranks on one NVLink/NVSwitch node): every rank allocates an inbox of P slices, exports it
// with CUDA IPC, and maps the inboxes of the others.  k_apply then stores the {hash, state} of a record straight
// into the owner's inbox, so that nothing is left to send when ptx_finalize runs.  Needs the expected record count
// (ptx_reserve before ptx_comm_init) on every rank; any failure leaves the ncclSend/ncclRecv path in place -
// decided collectively, so that all ranks take the same path.
int p2p_setup(ptx_ctx* ctx) {
    const int P = ctx->n_ranks;
    const int64_t test_cap = ctx->test_box_cap;
    ctx->test_box_cap = 0;
    int rc = xchg_ensure(ctx, 0);  // stream, events, cursors, box pointer array
    if (rc) return rc;
    unsigned long long* d_v = nullptr;
    uint8_t* d_h = nullptr;
    if ((rc = dalloc(ctx, &d_v, (size_t)P + 1)) || (rc = dalloc(ctx, &d_h, (size_t)(P + 1) * sizeof(cudaIpcMemHandle_t)))) return rc;
    std::vector<unsigned long long> v(P, 0);
    unsigned long long mine = (unsigned long long)std::max<int64_t>(ctx->reserve_records, 0);
    CU(cudaMemcpyAsync(d_v + P, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->st));
    if ((rc = nccl_check(ctx, g_nccl.AllGather(d_v + P, d_v, 1, ncclUint64, ctx->comm, ctx->st), "ncclAllGather(reserve)"))) return rc;
    CU(cudaMemcpyAsync(v.data(), d_v, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    unsigned long long lo = ~0ull, hi = 0;
    for (auto x : v) { lo = std::min(lo, x); hi = std::max(hi, x); }
    bool ok = lo > 0;  // every rank gave a size hint
    const uint64_t cap = !ok ? 0 : test_cap > 0 ? (uint64_t)test_cap : hi / P + hi / (4 * P) + 4096;
    cudaIpcMemHandle_t hmine;
    memset(&hmine, 0, sizeof hmine);
    if (ok) {
        ok = cudaMalloc((void**)&ctx->p2p_inbox, (size_t)P * cap * sizeof(ulonglong2)) == cudaSuccess &&
             cudaIpcGetMemHandle(&hmine, ctx->p2p_inbox) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    std::vector<cudaIpcMemHandle_t> hs(P);
    CU(cudaMemcpyAsync(d_h + (size_t)P * sizeof hmine, &hmine, sizeof hmine, cudaMemcpyHostToDevice, ctx->st));
    if ((rc = nccl_check(ctx, g_nccl.AllGather(d_h + (size_t)P * sizeof hmine, d_h, sizeof hmine, ncclUint8, ctx->comm, ctx->st), "ncclAllGather(ipc handles)"))) return rc;
    CU(cudaMemcpyAsync(hs.data(), d_h, (size_t)P * sizeof hmine, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->p2p_peer.assign(P, nullptr);
    std::vector<ulonglong2*> ptrs(P, nullptr);
    for (int q = 0; ok && q < P; ++q) {
        if (q == ctx->rank) continue;
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, hs[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->p2p_peer[q] = base;
        ptrs[q] = reinterpret_cast<ulonglong2*>(base) + (uint64_t)ctx->rank * cap;  // my slice of rank q's inbox
    }
    // all or nothing
    unsigned long long good = ok ? 1ull : 0ull;
    CU(cudaMemcpyAsync(d_v + P, &good, sizeof good, cudaMemcpyHostToDevice, ctx->st));
    if ((rc = nccl_check(ctx, g_nccl.AllReduce(d_v + P, d_v, 1, ncclUint64, ncclMin, ctx->comm, ctx->st), "ncclAllReduce(p2p ok)"))) return rc;
    CU(cudaMemcpyAsync(&good, d_v, sizeof good, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(d_v);
    cudaFree(d_h);
    if (!good) {
        for (void*& m : ctx->p2p_peer) { if (m) cudaIpcCloseMemHandle(m); m = nullptr; }
        dfree(ctx->p2p_inbox);
        return PTX_OK;
    }
    CU(cudaMemcpy(ctx->d_box_ptr, ptrs.data(), (size_t)P * sizeof(ulonglong2*), cudaMemcpyHostToDevice));
    ctx->p2p = true;
    ctx->p2p_cap = cap;
    ctx->box_cap = cap;
    return PTX_OK;
}

}  // namespace

extern "C" {

const char* ptx_version(void) { return "pantax_b200 0.1 (sm_100a)"; }

int ptx_create(int device, ptx_ctx** out) {
    if (!out) return PTX_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return PTX_E_CUDA;  // no CPU fallback by design
    if (cudaSetDevice(device) != cudaSuccess) return PTX_E_CUDA;
    if (const char* e = getenv("PTX_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));  // measurement knob: 32 / 64 / 128
    ptx_ctx* ctx = new (std::nothrow) ptx_ctx;
    if (!ctx) return PTX_E_NOMEM;
    ctx->device = device;
    if (const char* e = getenv("PTX_TILE_ROWS")) ctx->force_rows = atoi(e);
    if (const char* e = getenv("PTX_TILE_BYTES")) ctx->force_tile = atoi(e);
    if (const char* e = getenv("PTX_OLD_INGEST")) ctx->old_short = atoi(e) != 0;
    if (const char* e = getenv("PTX_LONG_NEW")) ctx->long_new = atoi(e) != 0;
    if (const char* e = getenv("PTX_LONG_TILE")) ctx->long_tile_max = std::min<uint32_t>(MAX_TILE, std::max<uint32_t>(MICRO, (uint32_t)atoi(e) / MICRO * MICRO));
    if (const char* e = getenv("PTX_NO_SORT")) ctx->no_sort = atoi(e) != 0;
    if (const char* e = getenv("PTX_KEEP_TEXT")) ctx->keep_text = atoi(e) != 0;
    if (const char* e = getenv("PTX_SCATTER")) ctx->scatter_var = atoi(e);
    if (const char* e = getenv("PTX_L2_HINTS")) ctx->l2_hints = atoi(e);
    if (const char* e = getenv("PTX_PF_WAVES")) ctx->pf_waves = atoi(e);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (const char* e = getenv("PTX_LONG_MODE")) ctx->force_long = atoi(e) ? 1 : 0;
    if (const char* e = getenv("PTX_TEST_BOX_CAP")) ctx->test_box_cap = atoll(e);
    if (const char* e = getenv("PTX_NO_SINGLE_PASS")) ctx->single_pass_ok = atoi(e) == 0;
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return PTX_E_CUDA;
    }
    if (cudaMalloc((void**)&ctx->d_flags, 4 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_total, sizeof(uint64_t)) != cudaSuccess) {
        delete ctx;
        return PTX_E_NOMEM;
    }
    cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(uint32_t), ctx->st);
    *out = ctx;
    return PTX_OK;
}

void ptx_destroy(ptx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& ch : ctx->chunks) chunk_free(ch);
    for (auto& ch : ctx->pool) chunk_free(ch);
    for (auto& t : ctx->text_pool) { if (t.buf) cudaFree(t.buf); if (t.copied) cudaEventDestroy(t.copied); }
    ctx->text_pool.clear();
    dfree(ctx->scratch);
    for (void* m : ctx->p2p_peer) if (m) cudaIpcCloseMemHandle(m);
    ctx->p2p_peer.clear();
    dfree(ctx->p2p_inbox); dfree(ctx->d_box_ptr);
    dfree(ctx->outbox); dfree(ctx->inbox); dfree(ctx->out_cursor);
    if (ctx->xs) cudaStreamDestroy(ctx->xs);
    if (ctx->ev_x0) cudaEventDestroy(ctx->ev_x0);
    if (ctx->ev_x1) cudaEventDestroy(ctx->ev_x1);
    free_graph(ctx);
    dfree(ctx->d_rstart); dfree(ctx->d_rend); dfree(ctx->d_node_base); dfree(ctx->d_order); dfree(ctx->d_sstart);
    dfree(ctx->d_hist); dfree(ctx->d_hist_g); dfree(ctx->d_flags); dfree(ctx->d_err); dfree(ctx->d_ds); dfree(ctx->d_pair_key); dfree(ctx->d_pair_val); dfree(ctx->d_pair_tmp); dfree(ctx->d_total); dfree(ctx->d_labels_in);
    ev_clear(ctx->ev_count);
    ev_clear(ctx->ev_ingest);
    ev_clear(ctx->ev_apply);
    ev_clear(ctx->ev_final);
    if (ctx->comm_x && ctx->comm_x != ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm_x);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    delete ctx;
}

const char* ptx_last_error(const ptx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int ptx_set_ranges(ptx_ctx* ctx, int S, const char* const* taxid, const int64_t* start, const int64_t* end) {
    if (!ctx || S <= 0 || !taxid || !start || !end) return fail(ctx, PTX_E_INVALID, "ptx_set_ranges: bad arguments");
    if (S >= 0xFFFFFD) return fail(ctx, PTX_E_UNSUPPORTED, "more than 2^24 - 3 species (the id-set slots keep the species in 24 bits)");
    if (!ctx->chunks.empty() || ctx->graphs_committed) return fail(ctx, PTX_E_STATE, "ranges must be set before graphs and GAF");
    cudaSetDevice(ctx->device);
    for (int s = 0; s < S; ++s)  // node ids are u32 in the reference too (profile.rs:547-551 SpeciesRange{start: u32, end: u32})
        if (end[s] > 0xFFFFFFFFll || start[s] < 0) return fail(ctx, PTX_E_UNSUPPORTED, "species range beyond 32-bit node ids");
    ctx->sp.assign(S, SpeciesHost());
    std::vector<uint32_t> order(S);
    for (int s = 0; s < S; ++s) {
        ctx->sp[s].taxid = taxid[s] ? taxid[s] : "";
        ctx->sp[s].start = start[s];
        ctx->sp[s].end = end[s];
        order[s] = (uint32_t)s;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return start[a] < start[b]; });
    ctx->disjoint = 1;
    for (int i = 0; i < S; ++i) {
        if (end[order[i]] < start[order[i]]) ctx->disjoint = 0;
        if (i + 1 < S && end[order[i]] >= start[order[i + 1]]) ctx->disjoint = 0;
    }
    dfree(ctx->d_rstart); dfree(ctx->d_rend); dfree(ctx->d_node_base); dfree(ctx->d_order); dfree(ctx->d_sstart); dfree(ctx->d_hist); dfree(ctx->d_err);
    ctx->labels_in_n = 0;  // supplied labels index the previous species list
    int rc;
    if ((rc = dalloc(ctx, &ctx->d_rstart, S)) || (rc = dalloc(ctx, &ctx->d_rend, S)) || (rc = dalloc(ctx, &ctx->d_node_base, S)) ||
        (rc = dalloc(ctx, &ctx->d_order, S)) || (rc = dalloc(ctx, &ctx->d_sstart, S)) || (rc = dalloc(ctx, &ctx->d_hist, (size_t)S * 4)) || (rc = dalloc(ctx, &ctx->d_err, S)))
        return rc;
    std::vector<int64_t> nb(S, -1);
    CU(cudaMemcpyAsync(ctx->d_rstart, start, S * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_rend, end, S * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_node_base, nb.data(), S * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_order, order.data(), S * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->st));
    std::vector<uint32_t> sstart(S);
    for (int i = 0; i < S; ++i) sstart[i] = (uint32_t)start[order[i]];
    CU(cudaMemcpyAsync(ctx->d_sstart, sstart.data(), S * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->h_err.assign(S, 0);
    return PTX_OK;
}

int ptx_upload_graph(ptx_ctx* ctx, int s, const int64_t* nodes_len, int64_t n, const uint64_t* path_off, const uint64_t* path_nodes,
                     int64_t H) {
    int rc = check_species(ctx, s, false);
    if (rc) return rc;
    if (ctx->graphs_committed) return fail(ctx, PTX_E_STATE, "graphs already committed");
    if (!nodes_len || n <= 0 || H < 0 || (H > 0 && (!path_off || !path_nodes))) return fail(ctx, PTX_E_INVALID, "ptx_upload_graph: bad arguments");
    SpeciesHost& sp = ctx->sp[s];
    if (sp.end - sp.start + 1 != n)  // profile.rs:2938: nvert = end - start + 1 indexes nodes_len
        return fail(ctx, PTX_E_NVERT_MISMATCH, "species %s: range %lld..%lld has %lld ids but the graph has %lld nodes", sp.taxid.c_str(),
                    (long long)sp.start, (long long)sp.end, (long long)(sp.end - sp.start + 1), (long long)n);
    sp.len.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        if (nodes_len[i] <= 0) return fail(ctx, PTX_E_ZERO_LEN, "species %s: node %lld has length %lld (profile.rs:494)", sp.taxid.c_str(), (long long)i, (long long)nodes_len[i]);
        if (nodes_len[i] > 0x7FFFFFFFll) return fail(ctx, PTX_E_INVALID, "node length exceeds 2^31");
        sp.len[i] = (uint32_t)nodes_len[i];
    }
    sp.path_off.assign(H + 1, 0);
    if (H > 0) {
        if (path_off[0] != 0) return fail(ctx, PTX_E_INVALID, "path_off[0] must be 0");
        for (int64_t h = 0; h <= H; ++h) {
            if (h && path_off[h] < path_off[h - 1]) return fail(ctx, PTX_E_INVALID, "path_off must be non-decreasing");
            sp.path_off[h] = path_off[h];
        }
    }
    const uint64_t P = H > 0 ? path_off[H] : 0;
    sp.path_nodes.resize(P);
    for (uint64_t k = 0; k < P; ++k) {
        if (path_nodes[k] >= (uint64_t)n) return fail(ctx, PTX_E_INVALID, "species %s: path node id %llu >= %lld nodes", sp.taxid.c_str(), (unsigned long long)path_nodes[k], (long long)n);
        sp.path_nodes[k] = (uint32_t)path_nodes[k];
    }
    sp.n_nodes = n;
    sp.n_paths = H;
    sp.uploaded = true;
    return PTX_OK;
}

// the per-species GFA (profile.rs:466-545 read_gfa, previous = 0) parsed on the device.  The text is uploaded once;
// kernels index its lines, read the S lines into node lengths and decode the path fields of the P / W lines node-parallel; only the
// compact arrays (4 bytes per node and per path step) and the list of path lines (their haplotype names) come back to the host,
// which lays the lines of one haplotype out behind each other (BTreeMap order of the names) and hands everything to the same
// staging ptx_upload_graph fills.
int ptx_upload_graph_gfa(ptx_ctx* ctx, int s, const uint8_t* gfa, size_t n) {
    int rc = check_species(ctx, s, false);
    if (rc) return rc;
    if (ctx->graphs_committed) return fail(ctx, PTX_E_STATE, "graphs already committed");
    if (!gfa || n == 0) return fail(ctx, PTX_E_INVALID, "ptx_upload_graph_gfa: bad arguments");
    cudaSetDevice(ctx->device);
    SpeciesHost& sp = ctx->sp[s];
    const int64_t n_nodes = sp.end - sp.start + 1;  // profile.rs:2938: nvert = end - start + 1 indexes nodes_len
    cudaStream_t st = ctx->st;
    Trace tr(st);
    const bool add_nl = gfa[n - 1] != '\n';
    const uint64_t nn = n + (add_nl ? 1 : 0);
    const uint32_t n_micro = (uint32_t)((nn + MICRO - 1) / MICRO);
    uint8_t* d_text = nullptr;
    uint32_t *d_cnt = nullptr, *d_len = nullptr, *d_is_s = nullptr, *d_adj = nullptr, *d_flags = nullptr, *d_pcnt = nullptr, *d_out = nullptr;
    uint64_t *d_base = nullptr, *d_scratch = nullptr, *d_line = nullptr, *d_ord = nullptr, *d_pieces = nullptr, *d_poff = nullptr;
    unsigned long long* d_np = nullptr;
    uint8_t* d_plist = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        cudaFree(d_text); cudaFree(d_cnt); cudaFree(d_len); cudaFree(d_is_s); cudaFree(d_adj); cudaFree(d_flags); cudaFree(d_pcnt); cudaFree(d_out);
        cudaFree(d_base); cudaFree(d_scratch); cudaFree(d_line); cudaFree(d_ord); cudaFree(d_pieces); cudaFree(d_poff); cudaFree(d_np); cudaFree(d_plist);
    };
#define GFA(call) do { if ((rc = (call)) != PTX_OK) { cleanup(); return rc; } } while (0)
#define GFA_SYNC(what) do { if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); return fail(ctx, PTX_E_CUDA, "ptx_upload_graph_gfa: %s: %s", what, cudaGetErrorString(cudaGetLastError())); } } while (0)
    GFA(dalloc(ctx, &d_text, (size_t)n_micro * MICRO + 64, false));
    if (cudaMemcpyAsync(d_text, gfa, n, cudaMemcpyHostToDevice, st) != cudaSuccess) { cleanup(); return fail(ctx, PTX_E_CUDA, "H2D copy failed"); }
    cudaMemsetAsync(d_text + n, '\n', (size_t)n_micro * MICRO + 64 - n, st);
    GFA(dalloc(ctx, &d_cnt, n_micro, false));
    GFA(dalloc(ctx, &d_base, (size_t)n_micro + 1, false));
    GFA(dalloc(ctx, &d_scratch, (size_t)n_micro / 2048 + 4, false));
    launch_flt_count_nl(d_text, nn, n_micro, d_cnt, st);
    launch_scan_u32(d_cnt, d_base, n_micro, d_scratch, st);
    uint64_t n_lines = 0;
    cudaMemcpyAsync(&n_lines, d_base + n_micro, sizeof n_lines, cudaMemcpyDeviceToHost, st);
    GFA_SYNC("newline index");
    tr.mark("gfa h2d + newline index");
    GFA(dalloc(ctx, &d_line, (size_t)n_lines + 2, false));
    launch_flt_line_starts(d_text, nn, n_micro, d_base, d_line, st);
    // ---- lines: node lengths, S-line order, list of path lines
    GFA(dalloc(ctx, &d_len, (size_t)n_nodes));
    GFA(dalloc(ctx, &d_is_s, (size_t)n_lines + 1, false));
    GFA(dalloc(ctx, &d_adj, (size_t)n_lines + 1, false));
    GFA(dalloc(ctx, &d_ord, (size_t)n_lines + 2, false));
    GFA(dalloc(ctx, &d_flags, 4));
    GFA(dalloc(ctx, &d_np, 1));
    uint64_t pcap = 4096;
    std::vector<GfaPathLineHost> plines;
    for (;;) {  // the path-line list is sized by guess and the pass repeated once if a GFA has more P / W lines than that
        GFA(dalloc(ctx, &d_plist, (size_t)pcap * gfa_path_line_bytes(), false));
        cudaMemsetAsync(d_np, 0, sizeof(unsigned long long), st);
        cudaMemsetAsync(d_flags, 0, 4 * sizeof(uint32_t), st);
        cudaMemsetAsync(d_len, 0, (size_t)n_nodes * sizeof(uint32_t), st);
        launch_gfa_lines(d_text, d_line, n_lines, n_nodes, d_len, d_is_s, d_adj, d_plist, d_np, pcap, d_flags, st);
        unsigned long long np = 0;
        cudaMemcpyAsync(&np, d_np, sizeof np, cudaMemcpyDeviceToHost, st);
        GFA_SYNC("line pass");
        if (np <= pcap) {
            plines.resize((size_t)np);
            if (np) cudaMemcpy(plines.data(), d_plist, (size_t)np * sizeof(GfaPathLineHost), cudaMemcpyDeviceToHost);
            break;
        }
        dfree(d_plist);
        pcap = np;
    }
    tr.mark("gfa line pass");
    dfree(d_scratch);
    GFA(dalloc(ctx, &d_scratch, (size_t)n_lines / 2048 + 4, false));
    launch_scan_u32(d_is_s, d_ord, n_lines, d_scratch, st);
    launch_gfa_check_order(d_is_s, d_adj, d_ord, n_lines, d_flags, st);
    uint64_t n_s = 0;
    uint32_t h_flags[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(&n_s, d_ord + n_lines, sizeof n_s, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
    GFA_SYNC("order check");
    auto flag_error = [&](uint32_t f) -> int {
        cleanup();
        if (f & 4u) return fail(ctx, PTX_E_ZERO_LEN, "species %s: node length 0 appears in the GFA (profile.rs:494)", sp.taxid.c_str());
        if (f & 2u) return fail(ctx, PTX_E_NODE_ORDER, "species %s: node ID out of order or mismatch (profile.rs:489)", sp.taxid.c_str());
        if (f & 1u) return fail(ctx, PTX_E_INVALID, "species %s: an S line's id is not a number or lies outside the species' %lld nodes", sp.taxid.c_str(), (long long)n_nodes);
        if (f & 8u) return fail(ctx, PTX_E_INVALID, "species %s: a P / W line has no name field", sp.taxid.c_str());
        return fail(ctx, PTX_E_INVALID, "species %s: a path names a node outside the graph", sp.taxid.c_str());
    };
    if (h_flags[0]) return flag_error(h_flags[0]);
    if ((int64_t)n_s != n_nodes) {
        cleanup();
        return fail(ctx, PTX_E_NVERT_MISMATCH, "species %s: range %lld..%lld has %lld ids but the GFA has %lld S lines", sp.taxid.c_str(), (long long)sp.start,
                    (long long)sp.end, (long long)n_nodes, (long long)n_s);
    }
    tr.mark("gfa order check");
    // ---- path lines in file order; names from the text (tiny copies); pieces of 4 KB over their path fields
    std::sort(plines.begin(), plines.end(), [](const GfaPathLineHost& a, const GfaPathLineHost& b) { return a.line < b.line; });
    std::vector<std::string> names(plines.size());
    for (size_t k = 0; k < plines.size(); ++k) {
        names[k].resize(plines[k].name_len);
        if (plines[k].name_len) cudaMemcpy(&names[k][0], d_text + plines[k].name_beg, plines[k].name_len, cudaMemcpyDeviceToHost);
    }
    std::vector<uint64_t> pbeg, pfend, pfbeg;
    std::vector<uint32_t> first_piece(plines.size() + 1, 0), pline;
    auto cut_pieces = [&]() {
        pbeg.clear(); pfend.clear(); pfbeg.clear(); pline.clear();
        for (size_t k = 0; k < plines.size(); ++k) {
            first_piece[k] = (uint32_t)pbeg.size();
            for (uint64_t p = plines[k].fld_beg; p < plines[k].fld_end; p += GFA_PIECE_BYTES) {
                pbeg.push_back(p); pfend.push_back(plines[k].fld_end); pfbeg.push_back(plines[k].fld_beg); pline.push_back((uint32_t)k);
            }
        }
        first_piece[plines.size()] = (uint32_t)pbeg.size();
    };
    cut_pieces();
    if (!pbeg.empty()) {  // where the path field really ends: the first tab behind its start (found in parallel; a path field is megabytes long)
        const uint32_t np0 = (uint32_t)pbeg.size();
        uint64_t* d_tmp = nullptr;
        uint32_t* d_pl = nullptr;
        unsigned long long* d_tab = nullptr;
        std::vector<unsigned long long> tab(plines.size(), ~0ull);
        if ((rc = dalloc(ctx, &d_tmp, (size_t)np0 * 2, false)) || (rc = dalloc(ctx, &d_pl, (size_t)np0, false)) || (rc = dalloc(ctx, &d_tab, plines.size(), false))) {
            cudaFree(d_tmp); cudaFree(d_pl); cudaFree(d_tab); cleanup(); return rc;
        }
        cudaMemcpyAsync(d_tmp, pbeg.data(), (size_t)np0 * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_tmp + np0, pfend.data(), (size_t)np0 * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_pl, pline.data(), (size_t)np0 * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_tab, tab.data(), plines.size() * 8, cudaMemcpyHostToDevice, st);
        launch_gfa_field_end(d_text, d_tmp, d_tmp + np0, d_pl, d_tab, np0, st);
        cudaMemcpyAsync(tab.data(), d_tab, plines.size() * 8, cudaMemcpyDeviceToHost, st);
        const bool ok = cudaStreamSynchronize(st) == cudaSuccess;
        cudaFree(d_tmp); cudaFree(d_pl); cudaFree(d_tab);
        if (!ok) { cleanup(); return fail(ctx, PTX_E_CUDA, "ptx_upload_graph_gfa: field ends: %s", cudaGetErrorString(cudaGetLastError())); }
        for (size_t k = 0; k < plines.size(); ++k) {
            if (tab[k] == ~0ull) continue;
            if (plines[k].kind == 2u) {
                cleanup();
                return fail(ctx, PTX_E_UNSUPPORTED, "species %s: a W line has fields behind its walk (the device parser takes the 7th field as the walk)", sp.taxid.c_str());
            }
            plines[k].fld_end = tab[k];
        }
        cut_pieces();
    }
    tr.mark("gfa names + field ends");
    const uint32_t n_pieces = (uint32_t)pbeg.size();
    std::vector<uint64_t> poff((size_t)n_pieces + 1, 0);
    if (n_pieces) {
        GFA(dalloc(ctx, &d_pieces, (size_t)n_pieces * 4, false));
        GFA(dalloc(ctx, &d_pcnt, (size_t)n_pieces, false));
        GFA(dalloc(ctx, &d_poff, (size_t)n_pieces + 2, false));
        cudaMemcpyAsync(d_pieces, pbeg.data(), (size_t)n_pieces * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_pieces + n_pieces, pfend.data(), (size_t)n_pieces * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_pieces + 2 * (size_t)n_pieces, pfbeg.data(), (size_t)n_pieces * 8, cudaMemcpyHostToDevice, st);
        launch_gfa_path_count(d_text, d_pieces, d_pieces + n_pieces, d_pcnt, n_pieces, d_flags, st);
        dfree(d_scratch);
        GFA(dalloc(ctx, &d_scratch, (size_t)n_pieces / 2048 + 4, false));
        launch_scan_u32(d_pcnt, d_poff, n_pieces, d_scratch, st);
        cudaMemcpyAsync(poff.data(), d_poff, ((size_t)n_pieces + 1) * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
        GFA_SYNC("path count");

    }
    tr.mark("gfa path count");
    // haplotypes in BTreeMap (byte) order of their names; the lines of one haplotype behind each other in file order
    std::map<std::string, std::vector<size_t>> haps;
    for (size_t k = 0; k < plines.size(); ++k) haps[names[k]].push_back(k);
    const uint64_t total_steps = poff[n_pieces];
    std::vector<uint64_t> path_off;
    std::vector<uint64_t> line_dst(plines.size(), 0);
    uint64_t run = 0;
    for (auto& kv : haps) {
        path_off.push_back(run);
        for (size_t k : kv.second) {
            line_dst[k] = run;
            run += poff[first_piece[k + 1]] - poff[first_piece[k]];
        }
    }
    path_off.push_back(run);
    std::vector<uint32_t> h_nodes((size_t)total_steps);
    if (n_pieces && total_steps) {
        std::vector<uint64_t> pdst((size_t)n_pieces);
        for (size_t k = 0; k < plines.size(); ++k)
            for (uint32_t c = first_piece[k]; c < first_piece[k + 1]; ++c) pdst[c] = line_dst[k] + (poff[c] - poff[first_piece[k]]);
        GFA(dalloc(ctx, &d_out, (size_t)total_steps, false));
        cudaMemcpyAsync(d_pieces + 3 * (size_t)n_pieces, pdst.data(), (size_t)n_pieces * 8, cudaMemcpyHostToDevice, st);
        launch_gfa_path_decode(d_text, d_pieces, d_pieces + n_pieces, d_pieces + 2 * (size_t)n_pieces, d_pieces + 3 * (size_t)n_pieces, n_nodes, d_out, n_pieces,
                               d_flags, st);
        cudaMemcpyAsync(h_nodes.data(), d_out, (size_t)total_steps * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
        GFA_SYNC("path decode");
        if (h_flags[0]) return flag_error(h_flags[0]);
    }
    tr.mark("gfa path decode + d2h");
    sp.len.resize((size_t)n_nodes);
    cudaMemcpy(sp.len.data(), d_len, (size_t)n_nodes * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cleanup();
#undef GFA
#undef GFA_SYNC
    sp.path_off = path_off;
    sp.path_nodes.swap(h_nodes);
    sp.path_names.clear();
    for (auto& kv : haps) sp.path_names.push_back(kv.first);
    sp.n_nodes = n_nodes;
    sp.n_paths = (int64_t)haps.size();
    sp.uploaded = true;
    return PTX_OK;
}

// The graph of a species as the library holds it (after ptx_upload_graph / ptx_upload_graph_gfa): what profile::read_gfa returned.
// Any of the output pointers may be null.  path_off has n_paths + 1 entries, path_nodes path_off[n_paths].
int ptx_species_graph(ptx_ctx* ctx, int s, int64_t* nodes_len, uint64_t* path_off, uint64_t* path_nodes) {
    int rc = check_species(ctx, s, false);
    if (rc) return rc;
    const SpeciesHost& sp = ctx->sp[s];
    if (!sp.uploaded && !sp.has_graph) return fail(ctx, PTX_E_NO_GRAPH, "species %s has no graph", sp.taxid.c_str());
    if (nodes_len) for (size_t i = 0; i < sp.len.size(); ++i) nodes_len[i] = (int64_t)sp.len[i];
    if (path_off) for (size_t i = 0; i < sp.path_off.size(); ++i) path_off[i] = sp.path_off[i];
    if (path_nodes) for (size_t i = 0; i < sp.path_nodes.size(); ++i) path_nodes[i] = (uint64_t)sp.path_nodes[i];
    return PTX_OK;
}
int64_t ptx_species_path_steps(const ptx_ctx* ctx, int s) {
    if (!ctx || s < 0 || s >= (int)ctx->sp.size()) return PTX_E_RANGE;
    return (int64_t)ctx->sp[s].path_nodes.size();
}
// Haplotype id of path h (ptx_upload_graph_gfa only; "" for graphs uploaded as arrays, whose names the caller has).  Returns the
// length of the name; at most cap - 1 bytes and a terminating 0 are written.
int ptx_species_path_name(ptx_ctx* ctx, int s, int64_t h, char* buf, size_t cap) {
    int rc = check_species(ctx, s, false);
    if (rc) return rc;
    const SpeciesHost& sp = ctx->sp[s];
    if (h < 0 || h >= sp.n_paths) return fail(ctx, PTX_E_RANGE, "path index out of range");
    const std::string name = (size_t)h < sp.path_names.size() ? sp.path_names[(size_t)h] : std::string();
    if (buf && cap) {
        const size_t k = std::min(name.size(), cap - 1);
        memcpy(buf, name.data(), k);
        buf[k] = 0;
    }
    return (int)name.size();
}

int ptx_commit_graphs(ptx_ctx* ctx) {
    if (!ctx) return PTX_E_INVALID;
    if (ctx->graphs_committed) return fail(ctx, PTX_E_STATE, "graphs already committed");
    cudaSetDevice(ctx->device);
    const auto t_commit0 = std::chrono::steady_clock::now();
    const int S = (int)ctx->sp.size();
    GraphDev& g = ctx->g;
    int64_t N = 0, Htot = 0, P = 0;
    for (auto& sp : ctx->sp)
        if (sp.uploaded) {
            sp.node_base = N;
            sp.hap_base = Htot;
            N += sp.n_nodes;
            Htot += sp.n_paths;
            P += (int64_t)sp.path_nodes.size();
        }
    if (N == 0) return fail(ctx, PTX_E_STATE, "no graph uploaded");
    if (N >= 0x7FFFFFFFll) return fail(ctx, PTX_E_UNSUPPORTED, "more than 2^31 nodes on one GPU");
    // ---- host-side concatenation (graph setup, outside the timed hot path)
    std::vector<uint32_t> len((size_t)N);
    std::vector<uint64_t> bit_off((size_t)N + 1);
    std::vector<uint32_t> pnode((size_t)P);
    std::vector<uint64_t> poff((size_t)Htot + 1);
    std::vector<int64_t> node_base(S, -1);
    uint64_t bits = 0;
    {
        int64_t h = 0;
        uint64_t k = 0;
        for (int s = 0; s < S; ++s) {
            SpeciesHost& sp = ctx->sp[s];
            if (!sp.uploaded) continue;
            node_base[s] = sp.node_base;
            for (int64_t i = 0; i < sp.n_nodes; ++i) {
                len[sp.node_base + i] = sp.len[i];
                bit_off[sp.node_base + i] = bits;
                bits += sp.len[i];
            }
            for (int64_t p = 0; p < sp.n_paths; ++p) {
                poff[h++] = k;
                for (uint64_t q = sp.path_off[p]; q < sp.path_off[p + 1]; ++q) pnode[k++] = (uint32_t)(sp.node_base + sp.path_nodes[q]);
            }
        }
        poff[Htot] = k;
        bit_off[N] = bits;
    }
    g.N = N; g.Htot = Htot; g.P = P;
    g.n_bit_words = (bits + 31) / 32 + 1;
    int rc;
    if ((rc = dalloc(ctx, &g.len, N, false)) || (rc = dalloc(ctx, &g.bit_off, N + 1, false)) || (rc = dalloc(ctx, &g.bases, N)) ||
        (rc = dalloc(ctx, &g.full, N)) || (rc = dalloc(ctx, &g.bits, g.n_bit_words + BITS_SLICE_PAD)) || (rc = dalloc(ctx, &g.cov, N)) ||
        (rc = dalloc(ctx, &g.pnode, P, false)) || (rc = dalloc(ctx, &g.poff, Htot + 1, false)) || (rc = dalloc(ctx, &g.path_len_sum, Htot)) ||
        (rc = dalloc(ctx, &g.path_cov_sum, Htot)) || (rc = dalloc(ctx, &g.hap_nz, Htot)) || (rc = dalloc(ctx, &g.trio_start, Htot + 1)) ||
        (rc = dalloc(ctx, &g.ninfo, N, false)))
        return rc;
    CU(cudaMemcpyAsync(g.len, len.data(), N * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(g.bit_off, bit_off.data(), (N + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
    launch_ninfo_build(g.len, g.bit_off, g.ninfo, N, ctx->st);
    if (P) CU(cudaMemcpyAsync(g.pnode, pnode.data(), P * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(g.poff, poff.data(), (Htot + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_node_base, node_base.data(), S * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));

    // ---- distinct-node marks per path: one bit per (path, node of its species), one launch for all paths
    uint64_t* d_pbm_off = nullptr;
    uint32_t* d_pbase = nullptr;
    uint32_t* d_pbm = nullptr;
    if (Htot > 0 && P > 0) {
        std::vector<uint64_t> pbm_off((size_t)Htot + 1, 0);
        std::vector<uint32_t> pbase((size_t)Htot, 0);
        int64_t h = 0;
        for (auto& sp : ctx->sp)
            if (sp.uploaded)
                for (int64_t p = 0; p < sp.n_paths; ++p, ++h) {
                    pbase[(size_t)h] = (uint32_t)sp.node_base;
                    pbm_off[(size_t)h + 1] = pbm_off[(size_t)h] + (uint64_t)((sp.n_nodes + 31) / 32);
                }
        if ((rc = dalloc(ctx, &d_pbm_off, (size_t)Htot + 1, false)) || (rc = dalloc(ctx, &d_pbase, (size_t)Htot, false)) ||
            (rc = dalloc(ctx, &d_pbm, (size_t)pbm_off[(size_t)Htot] + 1)))
            return rc;
        CU(cudaMemcpyAsync(d_pbm_off, pbm_off.data(), ((size_t)Htot + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_pbase, pbase.data(), (size_t)Htot * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->st));
        launch_mark_path_dups(g.pnode, g.poff, Htot, P, d_pbm_off, d_pbase, d_pbm, ctx->st);
        CU(cudaStreamSynchronize(ctx->st));  // the host vectors above must outlive the copies
    }
    launch_path_len_sum(g.pnode, g.poff, Htot, P, g.len, g.path_len_sum, ctx->st);

    // ---- unique trios (profile.rs:658-740)
    int64_t n_windows = 0;
    for (auto& sp : ctx->sp)
        if (sp.uploaded)
            for (int64_t p = 0; p < sp.n_paths; ++p) {
                uint64_t l = sp.path_off[p + 1] - sp.path_off[p];
                if (l >= 3) n_windows += (int64_t)l - 2;
            }
    g.T = 0;
    if (n_windows > 0) {
        const uint64_t kcap = 1ull << std::max<uint32_t>(10, log2_ceil((uint64_t)n_windows * 2));
        if (kcap > 0x80000000ull) return fail(ctx, PTX_E_UNSUPPORTED, "trio window table too large");
        uint4* keys = nullptr;
        uint32_t* cnt = nullptr;
        uint32_t* flag = nullptr;
        uint64_t* scan = nullptr;
        uint64_t* scratch = nullptr;
        CU(cudaMalloc((void**)&keys, kcap * sizeof(uint4)));
        CU(cudaMemsetAsync(keys, 0xFF, kcap * sizeof(uint4), ctx->st));
        if ((rc = dalloc(ctx, &cnt, kcap)) || (rc = dalloc(ctx, &flag, P, false)) || (rc = dalloc(ctx, &scan, P + 1, false)) ||
            (rc = dalloc(ctx, &scratch, P / 2048 + 4, false)))
            return rc;
        launch_trio_count(g.pnode, g.poff, Htot, P, keys, cnt, (uint32_t)(kcap - 1), ctx->st);
        launch_trio_flag(g.pnode, g.poff, Htot, P, keys, cnt, (uint32_t)(kcap - 1), flag, ctx->st);
        launch_scan_u32(flag, scan, (uint64_t)P, scratch, ctx->st);
        uint64_t T = 0;
        CU(cudaMemcpyAsync(&T, scan + P, sizeof T, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        g.T = (int64_t)T;
        if (T > 0) {
            const uint64_t tcap = 1ull << std::max<uint32_t>(10, log2_ceil(T * 2));
            if ((rc = dalloc(ctx, &g.trio_key, T * 3, false)) || (rc = dalloc(ctx, &g.trio_len, T, false)) ||
                (rc = dalloc(ctx, &g.trio_owner, T, false)) || (rc = dalloc(ctx, &g.trio_bases, T)))
                return rc;
            CU(cudaMalloc((void**)&g.tt, tcap * sizeof(uint4)));
            CU(cudaMemsetAsync(g.tt, 0xFF, tcap * sizeof(uint4), ctx->st));
            g.tt_mask = (uint32_t)(tcap - 1);
        }
        if (T > 0)
            launch_trio_emit(g.pnode, g.poff, Htot, P, flag, scan, g.len, g.trio_key, g.trio_len, g.trio_owner, g.tt, g.tt_mask, g.ninfo,
                             g.trio_start, ctx->st);
        CU(cudaStreamSynchronize(ctx->st));
        CU(cudaGetLastError());
        dfree(keys); dfree(cnt); dfree(flag); dfree(scan); dfree(scratch);
    }
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    dfree(d_pbm_off);
    dfree(d_pbase);
    dfree(d_pbm);
    // per-species trio slices: trios are ordered by (global hap, position) and haps are grouped by species
    std::vector<uint64_t> tstart((size_t)Htot + 1, 0);
    if (g.T > 0) CU(cudaMemcpy(tstart.data(), g.trio_start, (Htot + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    for (auto& sp : ctx->sp)
        if (sp.uploaded) {
            sp.trio_base = (int64_t)tstart[sp.hap_base];
            sp.n_trios = (int64_t)tstart[sp.hap_base + sp.n_paths] - sp.trio_base;
            sp.has_graph = true;
            sp.uploaded = false;
            std::vector<uint32_t>().swap(sp.len);
            std::vector<uint32_t>().swap(sp.path_nodes);
            // path_off is kept (tiny) for ptx_species_paths consumers
        }
    ctx->graphs_committed = true;
    ctx->commit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_commit0).count();
    for (auto& ch : ctx->chunks) ch.covered = false;
    if (!ctx->chunks.empty()) ctx->dirty = true;
    return PTX_OK;
}

int ptx_reserve(ptx_ctx* ctx, int64_t expected_records) {
    if (!ctx || expected_records < 0) return PTX_E_INVALID;
    ctx->reserve_records = expected_records;
    return PTX_OK;
}

int ptx_host_alloc(size_t bytes, void** out) {
    if (!out) return PTX_E_INVALID;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? PTX_OK : PTX_E_NOMEM;
}
int ptx_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? PTX_OK : PTX_E_CUDA; }

int ptx_ingest_gaf(ptx_ctx* ctx, const uint8_t* bytes, size_t n, int is_last) {
    if (!ctx || (!bytes && n)) return fail(ctx, PTX_E_INVALID, "ptx_ingest_gaf: bad arguments");
    if (ctx->sp.empty()) return fail(ctx, PTX_E_STATE, "call ptx_set_ranges first");
    cudaSetDevice(ctx->device);
    // usable part: up to and including the last '\n' (everything if is_last)
    size_t usable = n;
    if (!is_last) {
        const void* nl = n ? memrchr(bytes, '\n', n) : nullptr;
        usable = nl ? (size_t)((const uint8_t*)nl - bytes) + 1 : 0;
    }
    if (usable == 0 && !(is_last && !ctx->carry.empty())) {
        ctx->carry.insert(ctx->carry.end(), bytes, bytes + n);
        return PTX_OK;
    }
    // split into pieces of <= PIECE bytes at line boundaries so H2D of piece k+1 overlaps the kernels of piece k
    const size_t PIECE = 64ull << 20;
    size_t off = 0;
    bool first = true;
    std::vector<size_t> piece_idx;
    while (off < usable || (first && !ctx->carry.empty())) {
        size_t take = std::min(PIECE, usable - off);
        if (off + take < usable) {
            const void* nl = memrchr(bytes + off, '\n', take);
            if (nl) take = (size_t)((const uint8_t*)nl - (bytes + off)) + 1;
            else {  // a single line longer than PIECE: extend to its end
                const void* e = memchr(bytes + off + take, '\n', usable - off - take);
                take = e ? (size_t)((const uint8_t*)e - (bytes + off)) + 1 : usable - off;
            }
        }
        Chunk ch;
        const size_t c0 = first ? ctx->carry.size() : 0;
        int rc = chunk_alloc(ctx, ch, c0 + take);
        if (rc) return rc;
        if (c0) CU(cudaMemcpyAsync(ch.buf + PRE, ctx->carry.data(), c0, cudaMemcpyHostToDevice, ctx->copy_st));
        if (take) CU(cudaMemcpyAsync(ch.buf + PRE + c0, bytes + off, take, cudaMemcpyHostToDevice, ctx->copy_st));
        ch.n = c0 + take;
        ch.host_text = true;
        CU(cudaEventRecord(ch.copied, ctx->copy_st));
        ctx->chunks.push_back(ch);
        piece_idx.push_back(ctx->chunks.size() - 1);
        off += take;
        first = false;
    }
    // carry for the next call (the copy above read ctx->carry asynchronously from pageable memory: it is staged by the
    // runtime before cudaMemcpyAsync returns, so it is safe to overwrite now)
    ctx->carry.assign(bytes + usable, bytes + n);
    for (size_t pi : piece_idx) {
        Chunk& ch = ctx->chunks[pi];
        CU(cudaStreamWaitEvent(ctx->st, ch.copied, 0));
        int rc = chunk_process(ctx, ch);
        if (rc) return rc;
    }
    // the caller may reuse its buffer when this returns: wait for the copies (not for the kernels)
    if (!piece_idx.empty()) CU(cudaEventSynchronize(ctx->chunks[piece_idx.back()].copied));
    return PTX_OK;
}

// Strain-only resume (profile.rs:3367-3379): the species column comes from reads_classification.tsv, row-aligned
// with the GAF, instead of from the walk.  Appends to the rows supplied so far.
int ptx_ingest_labels(ptx_ctx* ctx, const uint32_t* labels, int64_t n) {
    if (!ctx || (!labels && n) || n < 0) return fail(ctx, PTX_E_INVALID, "ptx_ingest_labels: bad arguments");
    if (ctx->sp.empty()) return fail(ctx, PTX_E_STATE, "call ptx_set_ranges first");
    if (ctx->total_records > ctx->labels_in_n) return fail(ctx, PTX_E_STATE, "ptx_ingest_labels: GAF rows were already ingested without labels");
    const uint32_t S = (uint32_t)ctx->sp.size();
    for (int64_t i = 0; i < n; ++i)
        if (labels[i] != LABEL_U && labels[i] >= S) return fail(ctx, PTX_E_RANGE, "ptx_ingest_labels: label is neither a species index nor PTX_LABEL_U");
    if (n == 0) return PTX_OK;
    cudaSetDevice(ctx->device);
    if (ctx->labels_in_n + n > ctx->labels_in_cap) {
        const int64_t cap = std::max<int64_t>(ctx->labels_in_n + n, 2 * ctx->labels_in_cap);
        uint32_t* nd = nullptr;
        CU(cudaMalloc((void**)&nd, (size_t)cap * sizeof(uint32_t)));
        CU(cudaStreamSynchronize(ctx->st));
        if (ctx->labels_in_n) CU(cudaMemcpy(nd, ctx->d_labels_in, (size_t)ctx->labels_in_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        dfree(ctx->d_labels_in);
        ctx->d_labels_in = nd;
        ctx->labels_in_cap = cap;
    }
    CU(cudaMemcpy(ctx->d_labels_in + ctx->labels_in_n, labels, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->labels_in_n += n;
    return PTX_OK;
}

int ptx_gaf_buffer_alloc(ptx_ctx* ctx, size_t capacity, int* buffer_id, void** device_ptr) {
    if (!ctx || !buffer_id || !device_ptr) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    Chunk ch;
    int rc = chunk_alloc(ctx, ch, capacity);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->copy_st));  // the newline padding in front of the text is in place before the caller fills the buffer
    ctx->chunks.push_back(ch);
    *buffer_id = (int)ctx->chunks.size() - 1;
    *device_ptr = ch.buf + PRE;
    return PTX_OK;
}

int ptx_ingest_gaf_device(ptx_ctx* ctx, int buffer_id, size_t n) {
    if (!ctx || buffer_id < 0 || buffer_id >= (int)ctx->chunks.size()) return fail(ctx, PTX_E_RANGE, "bad buffer id");
    Chunk& ch = ctx->chunks[buffer_id];
    if (ch.ingested) return fail(ctx, PTX_E_STATE, "buffer already ingested");
    if (n > ch.cap) return fail(ctx, PTX_E_INVALID, "n exceeds the buffer capacity");
    if (ctx->sp.empty()) return fail(ctx, PTX_E_STATE, "call ptx_set_ranges first");
    cudaSetDevice(ctx->device);
    ch.n = n;
    return chunk_process(ctx, ch);
}

int ptx_finalize(ptx_ctx* ctx) {
    if (!ctx) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!ctx->carry.empty()) return fail(ctx, PTX_E_STATE, "a partial line is pending: pass is_last=1 on the final chunk");
    for (auto& ch : ctx->chunks)
        if (!ch.ingested && ch.n == 0) { ch.ingested = true; ch.covered = true; }  // unused device buffers
    if (!ctx->dirty) return PTX_OK;
    GraphDev& g = ctx->g;
    const int S = (int)ctx->sp.size();
    {
        int rc = chunks_resolve(ctx);  // counts of the single-pass chunks (redoes the ones whose estimate was too small)
        if (rc) return rc;
    }
    ev_begin(ctx, ctx->ev_final);
    Trace tr(ctx->st);
    bool mixed = false;
    auto cover_pending = [&](bool keepmask) -> int {
        for (auto& ch : ctx->chunks) {
            if (!ch.ingested || ch.covered || ch.n_tiles == 0) continue;
            uint32_t n_nodes = 0;
            if (pick_scatter_variant(ctx) == 3) {  // (node, bases) pairs beside the CSR walks, sorted and reduced per chunk
                CU(cudaMemcpyAsync(&n_nodes, reinterpret_cast<uint32_t*>(ch.cursors) + 1, sizeof n_nodes, cudaMemcpyDeviceToHost, ctx->st));
                CU(cudaStreamSynchronize(ctx->st));
                const size_t tb = scatter_sorted_tmp_bytes(n_nodes);
                if (ctx->pair_cap < (int64_t)n_nodes || ctx->pair_tmp_cap < tb) {
                    dfree(ctx->d_pair_key); dfree(ctx->d_pair_val); dfree(ctx->d_pair_tmp);
                    ctx->pair_cap = (int64_t)n_nodes + n_nodes / 8 + 1024;
                    ctx->pair_tmp_cap = scatter_sorted_tmp_bytes((uint64_t)ctx->pair_cap);
                    CU(cudaMalloc((void**)&ctx->d_pair_key, (size_t)ctx->pair_cap * sizeof(uint32_t)));
                    CU(cudaMalloc((void**)&ctx->d_pair_val, (size_t)ctx->pair_cap * sizeof(unsigned long long)));
                    CU(cudaMalloc((void**)&ctx->d_pair_tmp, ctx->pair_tmp_cap));
                }
                CU(cudaMemsetAsync(ctx->d_pair_key, 0xFF, (size_t)n_nodes * sizeof(uint32_t), ctx->st));  // slots of records that are not covered
            }
            IngestArgs a = make_args(ctx, ch);
            launch_apply(a, (uint32_t)ch.n_slots, MODE_COVER | (keepmask ? MODE_KEEPMASK : 0), ctx->st);  // from the record table: no text is re-read
            if (a.scatter_var == 3) launch_scatter_sorted(ctx->d_pair_key, ctx->d_pair_val, n_nodes, ctx->g.bases, ctx->d_pair_tmp, ctx->pair_tmp_cap, ctx->st);
            ch.covered = true;
        }
        return PTX_OK;
    };
    if (ctx->comm) {
        // id groups may span ranks: the boxes k_apply filled travel on the side stream while the coverage runs here
        int rc = xchg_ensure(ctx, std::max<int64_t>(ctx->ds_records, ctx->reserve_records));
        if (rc) return rc;
        if (ctx->graphs_committed && g.N > 0 && ctx->cov_reduced) return fail(ctx, PTX_E_STATE, "multi-GPU: coverage already reduced");
        rc = exchange_begin(ctx, [&]() -> int { return (ctx->graphs_committed && g.N > 0) ? cover_pending(false) : PTX_OK; });
        if (rc) return rc;
        tr.mark("final exchange issued");
    } else {
        CU(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof ctx->h_flags, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        mixed = ctx->h_flags[1] != 0;
        tr.mark("final flags");
    }
    if (ctx->graphs_committed && g.N > 0) {
        auto start_over = [&]() -> int {
            // profile.rs:406-437: some id group spans species -> its reads must not contribute.  The optimistic
            // pass counted them: zero the accumulators and replay the record table with the keep mask.
            bool any_optimistic = false;
            for (auto& ch : ctx->chunks) any_optimistic |= (ch.ingested && ch.covered);
            if (any_optimistic) {
                int rc = zero_coverage(ctx);
                if (rc) return rc;
                for (auto& ch : ctx->chunks) ch.covered = false;
            }
            return PTX_OK;
        };
        auto reduce_and_stats = [&]() -> int {
            launch_ninfo_full(g.ninfo, g.full, g.N, 0, ctx->st);
            if (ctx->comm) {
                // int64 sums and ORs are order-free: bit-exact for any shard count.  All sums travel as ONE aggregated NCCL
                // launch (grouped all-reduces); the covered-base bitmap - the full-node flags written into it first - is
                // OR-reduced by slices: every rank receives the other ranks' copies of ITS slice (grouped send/recv), ORs
                // them, and the reduced slices are all-gathered in place (NCCL has no bitwise-or reduction).
                int rc;
                const int P = ctx->n_ranks;
                launch_bits_fill_full(g, ctx->st);
                if (!ctx->d_hist_g) { if ((rc = dalloc(ctx, &ctx->d_hist_g, (size_t)S * 4))) return rc; }
                const uint64_t slice = (((g.n_bit_words + P - 1) / P) + 3) & ~(uint64_t)3;  // P * slice <= n_bit_words + BITS_SLICE_PAD
                if ((rc = scratch_reserve(ctx, (size_t)(P - 1) * slice * sizeof(uint32_t) + 4096))) return rc;
                ctx->scratch_off = 0;
                uint32_t* got = scratch_take<uint32_t>(ctx, (size_t)(P - 1) * slice);
                g_nccl.GroupStart();
                g_nccl.AllReduce(g.bases, g.bases, g.N, ncclUint64, ncclSum, ctx->comm, ctx->st);
                if (g.T > 0) g_nccl.AllReduce(g.trio_bases, g.trio_bases, g.T, ncclUint64, ncclSum, ctx->comm, ctx->st);
                g_nccl.AllReduce(ctx->d_err, ctx->d_err, S, ncclUint32, ncclMax, ctx->comm, ctx->st);
                g_nccl.AllReduce(ctx->d_hist, ctx->d_hist_g, (size_t)S * 4, ncclUint64, ncclSum, ctx->comm, ctx->st);
                if ((rc = nccl_check(ctx, g_nccl.GroupEnd(), "grouped ncclAllReduce(bases, trio_bases, err, hist)"))) return rc;
                g_nccl.GroupStart();
                for (int q = 0, k = 0; q < P; ++q) {
                    if (q == ctx->rank) continue;
                    g_nccl.Send(g.bits + (uint64_t)q * slice, slice, ncclUint32, q, ctx->comm, ctx->st);
                    g_nccl.Recv(got + (uint64_t)k * slice, slice, ncclUint32, q, ctx->comm, ctx->st);
                    ++k;
                }
                if ((rc = nccl_check(ctx, g_nccl.GroupEnd(), "grouped ncclSend/ncclRecv(bitmap slices)"))) return rc;
                launch_or_slices(g.bits + (uint64_t)ctx->rank * slice, got, (uint32_t)(P - 1), slice, ctx->st);
                if ((rc = nccl_check(ctx, g_nccl.AllGather(g.bits + (uint64_t)ctx->rank * slice, g.bits, slice, ncclUint32, ctx->comm, ctx->st), "ncclAllGather(bitmap slices)"))) return rc;
                ctx->hist_reduced = true;
            }
            tr.mark("final reductions");
            CU(cudaMemsetAsync(g.path_cov_sum, 0, std::max<int64_t>(g.Htot, 1) * sizeof(unsigned long long), ctx->st));
            CU(cudaMemsetAsync(g.hap_nz, 0, std::max<int64_t>(g.Htot, 1) * sizeof(unsigned long long), ctx->st));
            launch_cov(g, ctx->comm != nullptr, ctx->st);
            launch_path_cov_sum(g, ctx->st);
            launch_hap_nz(g, ctx->st);
            return PTX_OK;
        };
        if (ctx->comm && ctx->cov_reduced) return fail(ctx, PTX_E_STATE, "multi-GPU: coverage already reduced");
        if (mixed) { int rc = start_over(); if (rc) return rc; }
        { int rc = cover_pending(mixed); if (rc) return rc; }
        bool reduced = false;
        if (ctx->comm && ctx->comm_x != ctx->comm) {
            // the exchange runs on its own communicator and stream: reduce optimistically (no mixed id group is
            // the common case) instead of idling until the owner-side merge has finished
            int rc = reduce_and_stats();
            if (rc) return rc;
            reduced = true;
        }
        if (ctx->comm) {
            CU(cudaStreamWaitEvent(ctx->st, ctx->ev_x1, 0));
            CU(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof ctx->h_flags, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
            if (ctx->h_flags[1]) {
                int rc = return_mixed_ids(ctx, ctx->st);
                if (rc) return rc;
            }
            mixed = ctx->h_flags[1] != 0;
            if (mixed) {
                int rc = start_over();
                if (rc) return rc;
                if ((rc = cover_pending(true))) return rc;
                reduced = false;
            }
        }
        tr.mark("final replay/cover");
        if (!reduced) {
            int rc = reduce_and_stats();
            if (rc) return rc;
        }
        if (ctx->comm) ctx->cov_reduced = true;
    }
    if (ctx->comm && !(ctx->graphs_committed && g.N > 0)) {  // no coverage pass ran: join the exchange here
        CU(cudaStreamWaitEvent(ctx->st, ctx->ev_x1, 0));
        CU(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof ctx->h_flags, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        if (ctx->h_flags[1]) { int rc = return_mixed_ids(ctx, ctx->st); if (rc) return rc; }
    }
    if (ctx->comm && !ctx->hist_reduced) {  // no coverage reduction ran (no graphs): the species counts travel alone
        if (!ctx->d_hist_g) { int rc = dalloc(ctx, &ctx->d_hist_g, (size_t)S * 4); if (rc) return rc; }
        int rc = nccl_check(ctx, g_nccl.AllReduce(ctx->d_hist, ctx->d_hist_g, (size_t)S * 4, ncclUint64, ncclSum, ctx->comm, ctx->st), "ncclAllReduce(hist)");
        if (rc) return rc;
    }
    ctx->hist_reduced = false;
    tr.mark("final cov/path/hap stats");
    ctx->h_err.assign(S, 0);
    CU(cudaMemcpyAsync(ctx->h_err.data(), ctx->d_err, S * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->st));
    ev_end(ctx, ctx->ev_final);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    ctx->dirty = false;
    if (ctx->labels_in_n > 0) {
        uint32_t bad = 0;
        CU(cudaMemcpy(&bad, ctx->d_flags + 3, sizeof bad, cudaMemcpyDeviceToHost));
        if (bad) return fail(ctx, PTX_E_RANGE, "ptx_ingest_labels: a row is labelled with a species whose node range does not contain its walk "
                                               "(such rows were treated as unclassified; the reference indexes out of the species graph here)");
    }
    return PTX_OK;
}

static int reset_impl(ptx_ctx* ctx, bool keep_buffers) {
    if (!ctx) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaStreamSynchronize(ctx->copy_st));
    ctx->pending_records = 0;
    for (auto& ch : ctx->chunks) ch.pending = false;
    if (keep_buffers) {
        for (auto& ch : ctx->chunks) { ch.ingested = false; ch.covered = false; ch.n_records = 0; }
    } else {
        size_t pooled = 0;
        for (auto& c : ctx->pool) pooled += c.cap;
        for (auto& ch : ctx->chunks) {
            if (pooled + ch.cap <= ((size_t)48 << 30) && ctx->pool.size() < 4096) { pooled += ch.cap; ctx->pool.push_back(ch); }
            else chunk_free(ch);
        }
        ctx->chunks.clear();
    }
    ctx->carry.clear();
    if (!keep_buffers) ctx->labels_in_n = 0;  // supplied labels belong to the input that is being dropped
    ctx->total_records = 0;
    ctx->labelled_rows_known = 0;
    ctx->first_rows.clear();
    ctx->first_rows_global = false;
    ctx->ds_records = 0;
    ctx->cov_reduced = false;
    ctx->h_flags[0] = ctx->h_flags[1] = ctx->h_flags[2] = 0;
    const size_t S = std::max<size_t>(ctx->sp.size(), 1);
    if (ctx->d_hist) CU(cudaMemsetAsync(ctx->d_hist, 0, S * 4 * sizeof(unsigned long long), ctx->st));
    CU(cudaMemsetAsync(ctx->d_flags, 0, 4 * sizeof(uint32_t), ctx->st));
    if (ctx->out_cursor) {
        CU(cudaMemsetAsync(ctx->out_cursor, 0, ((size_t)ctx->n_ranks + 1) * sizeof(unsigned long long), ctx->st));
        std::fill(ctx->box_sent.begin(), ctx->box_sent.end(), 0ull);
        std::fill(ctx->recv_done.begin(), ctx->recv_done.end(), 0ull);
    }
    ctx->ds_entries_bound = 0;
    { int rc = ds_new_pass(ctx); if (rc) return rc; }
    if (ctx->d_err) {
        int rc = zero_coverage(ctx);
        if (rc) return rc;
        CU(cudaMemsetAsync(ctx->d_err, 0, S * sizeof(uint32_t), ctx->st));
    }
    ctx->h_err.assign(ctx->sp.size(), 0);
    ev_clear(ctx->ev_count);
    ev_clear(ctx->ev_ingest);
    ev_clear(ctx->ev_apply);
    ev_clear(ctx->ev_final);
    ctx->dirty = false;
    return PTX_OK;
}

int ptx_reset(ptx_ctx* ctx) {
    int rc = reset_impl(ctx, false);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->st));
    return PTX_OK;
}

// Zero every accumulator (stream-ordered, no host sync) but keep the device GAF buffers, which
// can then be ingested again with ptx_ingest_gaf_device: one more pass over the same resident text.
int ptx_rewind(ptx_ctx* ctx) { return reset_impl(ctx, true); }

int64_t ptx_num_records(const ptx_ctx* ctx) {
    if (!ctx) return PTX_E_INVALID;
    if (chunks_resolve(const_cast<ptx_ctx*>(ctx)) != PTX_OK) return PTX_E_CUDA;
    return ctx->total_records;
}
int ptx_num_species(const ptx_ctx* ctx) { return ctx ? (int)ctx->sp.size() : PTX_E_INVALID; }
int ptx_ids_unique(const ptx_ctx* ctx) { return ctx ? (ctx->h_flags[0] == 0 ? 1 : 0) : PTX_E_INVALID; }

int ptx_read_labels(ptx_ctx* ctx, uint32_t* labels) {
    if (!ctx || !labels) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    {
        int rc = chunks_resolve(ctx);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->st));
    int64_t off = 0;
    for (auto& ch : ctx->chunks) {
        if (!ch.ingested || ch.n_records == 0) continue;
        int rc = chunk_labels_materialize(ctx, ch);
        if (rc) return rc;
        CU(cudaMemcpy(labels + off, ch.labels, ch.n_records * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        off += ch.n_records;
    }
    return PTX_OK;
}

int ptx_species_counts(ptx_ctx* ctx, int64_t* counts) {
    if (!ctx || !counts) return PTX_E_INVALID;
    if (ctx->sp.empty()) return fail(ctx, PTX_E_STATE, "no ranges");
    if (ctx->dirty && ctx->comm) return fail(ctx, PTX_E_STATE, "call ptx_finalize first");
    cudaSetDevice(ctx->device);
    {
        int rc = chunks_resolve(ctx);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaMemcpy(counts, ctx->comm ? ctx->d_hist_g : ctx->d_hist, ctx->sp.size() * 4 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return PTX_OK;
}

}  // extern "C"
namespace {
// read_len (column 2; NULL_I64 if null) of the first <= 1000 non-U rows this context ingested, in GAF order (profile.rs:311-322):
// on the host, from the first ~1000 lines of the first chunk(s)
int local_first_rows(ptx_ctx* ctx, std::vector<int64_t>& vals) {
    vals.clear();
    {
        int rc = chunks_resolve(ctx);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->st));
    int64_t seen = 0;
    for (auto& ch : ctx->chunks) {
        if (seen >= 1000) break;
        if (!ch.ingested || ch.n_records == 0) continue;
        if (!ch.buf) return fail(ctx, PTX_E_STATE, "ptx_equal_length: the text of the first rows is gone");  // cannot happen: text is kept until 1000 non-U rows are known
        {
            int rc = chunk_labels_materialize(ctx, ch);
            if (rc) return rc;
        }
        size_t text_off = 0, win = 1u << 20;
        int64_t rec = 0;
        std::vector<uint8_t> text;
        std::vector<uint32_t> lab;
        while (seen < 1000 && rec < ch.n_records && text_off < ch.n) {
            const size_t want = std::min<size_t>(ch.n - text_off, win);
            const bool to_end = text_off + want >= ch.n;
            text.resize(want + 2);
            text[want + 1] = '\n';
            CU(cudaMemcpy(text.data(), ch.buf + PRE + text_off, want, cudaMemcpyDeviceToHost));
            text[want] = '\n';  // terminates an unterminated last line when the window reaches the end of the chunk
            std::vector<std::pair<size_t, size_t>> lines;  // start, length incl. '\n'
            size_t i = 0, consumed = 0;
            while (i < want) {
                const void* nl = memchr(text.data() + i, '\n', want - i);
                if (!nl && !to_end) break;  // partial line: refetch from its start
                const size_t e = nl ? (size_t)((const uint8_t*)nl - text.data()) : want;
                size_t l = e - i;
                if (l && text[i + l - 1] == '\r') --l;
                if (l && text[i] != '@') lines.push_back({i, e - i + 1});
                i = e + 1;
                consumed = std::min(i, want);
            }
            if (consumed == 0) { win *= 2; continue; }  // one line longer than the window
            lab.resize(lines.size());
            if (!lines.empty()) CU(cudaMemcpy(lab.data(), ch.labels + rec, lines.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            for (size_t k = 0; k < lines.size() && seen < 1000; ++k) {
                if (lab[k] == LABEL_U) continue;
                RecParse r;
                parse_record(text.data() + lines[k].first, 0, 0xFFFFFFF0u, r, 1u, nullptr, 0, 0);  // the line's own '\n' stops the scanners;  // same column rules as the kernel
                vals.push_back(r.qlen);
                ++seen;
            }
            rec += (int64_t)lines.size();
            text_off += consumed;
        }
    }
    return PTX_OK;
}

// multi-GPU, inside ptx_finalize (a collective): the first 1000 non-U rows of the GAF are rank 0's - unless rank 0's read batch holds
// fewer; then the ranks behind it supply the rest, in rank order.  `counts[q]` = min(non-U rows of rank q, 1000), known to every
// rank from the fills all-gather, so all ranks take this branch together.
int first_rows_exchange(ptx_ctx* ctx, const std::vector<unsigned long long>& counts) {
    const int P = ctx->n_ranks;
    const size_t SLOT = 1001;  // count + 1000 values
    std::vector<int64_t> mine;
    int rc = local_first_rows(ctx, mine);
    if (rc) return rc;
    std::vector<int64_t> host((size_t)(P + 1) * SLOT, 0);
    host[(size_t)P * SLOT] = (int64_t)mine.size();
    std::copy(mine.begin(), mine.end(), host.begin() + (size_t)P * SLOT + 1);
    int64_t* d = nullptr;
    if ((rc = dalloc(ctx, &d, (size_t)(P + 1) * SLOT, false))) return rc;
    CU(cudaMemcpyAsync(d + (size_t)P * SLOT, host.data() + (size_t)P * SLOT, SLOT * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    if ((rc = nccl_check(ctx, g_nccl.AllGather(d + (size_t)P * SLOT, d, SLOT, ncclUint64, ctx->comm, ctx->st), "ncclAllGather(first rows)"))) { cudaFree(d); return rc; }
    CU(cudaMemcpyAsync(host.data(), d, (size_t)P * SLOT * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(d);
    ctx->first_rows.clear();
    for (int q = 0; q < P && ctx->first_rows.size() < 1000; ++q) {
        const int64_t n = std::min<int64_t>(host[(size_t)q * SLOT], 1000);
        (void)counts;
        for (int64_t k = 0; k < n && ctx->first_rows.size() < 1000; ++k) ctx->first_rows.push_back(host[(size_t)q * SLOT + 1 + (size_t)k]);
    }
    ctx->first_rows_global = true;
    return PTX_OK;
}
}  // namespace
extern "C" {

// profile.rs:311-322: are the read lengths of the first 1000 non-U rows all equal?
int ptx_equal_length(ptx_ctx* ctx, int* is_equal, int64_t* read_len) {
    if (!ctx || !is_equal || !read_len) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    std::vector<int64_t> vals;
    if (ctx->first_rows_global) vals = ctx->first_rows;  // multi-GPU: gathered over the ranks by ptx_finalize
    else {
        int rc = local_first_rows(ctx, vals);
        if (rc) return rc;
    }
    std::vector<int64_t> distinct;  // a null read_len (NULL_I64) counts as a value, as in polars' unique()
    for (int64_t v : vals)
        if (std::find(distinct.begin(), distinct.end(), v) == distinct.end()) distinct.push_back(v);
    *is_equal = distinct.size() == 1 ? 1 : 0;
    *read_len = distinct.size() == 1 ? distinct[0] : 0;
    return PTX_OK;
}

int64_t ptx_species_nodes(const ptx_ctx* ctx, int s) {
    if (!ctx || s < 0 || s >= (int)ctx->sp.size()) return PTX_E_RANGE;
    return ctx->sp[s].has_graph || ctx->sp[s].uploaded ? ctx->sp[s].n_nodes : PTX_E_NO_GRAPH;
}
int64_t ptx_species_paths(const ptx_ctx* ctx, int s) {
    if (!ctx || s < 0 || s >= (int)ctx->sp.size()) return PTX_E_RANGE;
    return ctx->sp[s].has_graph || ctx->sp[s].uploaded ? ctx->sp[s].n_paths : PTX_E_NO_GRAPH;
}
int64_t ptx_species_trios(const ptx_ctx* ctx, int s) {
    if (!ctx || s < 0 || s >= (int)ctx->sp.size()) return PTX_E_RANGE;
    return ctx->sp[s].has_graph ? ctx->sp[s].n_trios : PTX_E_NO_GRAPH;
}

int ptx_node_bases(ptx_ctx* ctx, int s, int64_t* out) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    if (!out) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    CU(cudaMemcpy(out, ctx->g.bases + sp.node_base, sp.n_nodes * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return PTX_OK;
}

int ptx_node_cov(ptx_ctx* ctx, int s, uint64_t* out) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    if (!out) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    std::vector<uint32_t> tmp((size_t)sp.n_nodes);
    CU(cudaMemcpy(tmp.data(), ctx->g.cov + sp.node_base, sp.n_nodes * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < sp.n_nodes; ++i) out[i] = tmp[i];
    return PTX_OK;
}

int ptx_node_depth(ptx_ctx* ctx, int s, double* out) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    if (!out) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, sp.n_nodes * sizeof(double)));
    launch_depth(ctx->g.bases + sp.node_base, ctx->g.len + sp.node_base, nullptr, d, (uint64_t)sp.n_nodes, ctx->st);
    CU(cudaMemcpyAsync(out, d, sp.n_nodes * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(d);
    return PTX_OK;
}

int ptx_trio_bases(ptx_ctx* ctx, int s, int64_t* out) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    if (sp.n_trios == 0) return PTX_OK;
    if (!out) return PTX_E_INVALID;
    CU(cudaMemcpy(out, ctx->g.trio_bases + sp.trio_base, sp.n_trios * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return PTX_OK;
}

int ptx_trio_depth(ptx_ctx* ctx, int s, double* out) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    if (sp.n_trios == 0) return PTX_OK;
    if (!out) return PTX_E_INVALID;
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, sp.n_trios * sizeof(double)));
    launch_depth(ctx->g.trio_bases + sp.trio_base, nullptr, ctx->g.trio_len + sp.trio_base, d, (uint64_t)sp.n_trios, ctx->st);
    CU(cudaMemcpyAsync(out, d, sp.n_trios * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    cudaFree(d);
    return PTX_OK;
}

int ptx_trio_table(ptx_ctx* ctx, int s, uint64_t* keys3, int64_t* len, uint32_t* owner) {
    int rc = check_species(ctx, s, true);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    const int64_t T = sp.n_trios;
    if (T == 0) return PTX_OK;
    if (keys3) {
        std::vector<uint32_t> k((size_t)T * 3);
        CU(cudaMemcpy(k.data(), ctx->g.trio_key + 3 * sp.trio_base, T * 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < T * 3; ++i) keys3[i] = (uint64_t)k[i] - (uint64_t)sp.node_base;
    }
    if (len) CU(cudaMemcpy(len, ctx->g.trio_len + sp.trio_base, T * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (owner) {
        CU(cudaMemcpy(owner, ctx->g.trio_owner + sp.trio_base, T * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < T; ++i) owner[i] -= (uint32_t)sp.hap_base;
    }
    return PTX_OK;
}

int ptx_trio_ref_order(const uint64_t* path_off, const uint64_t* path_nodes, int64_t n_paths, const uint64_t* keys3, int64_t n_trios,
                       uint64_t* order) {
    if (n_paths < 0 || n_trios < 0 || (n_paths > 0 && (!path_off || (path_off[n_paths] > 0 && !path_nodes))) ||
        (n_trios > 0 && (!keys3 || !order)))
        return PTX_E_INVALID;
    try {
        return ptx_fx::trio_ref_order(path_off, path_nodes, n_paths, keys3, n_trios, order) == 0 ? PTX_OK : PTX_E_INVALID;
    } catch (const std::bad_alloc&) {
        return PTX_E_NOMEM;
    }
}

int ptx_path_sums(ptx_ctx* ctx, int s, int64_t* sum_cov, int64_t* sum_len) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    if (sp.n_paths == 0) return PTX_OK;
    if (sum_cov) CU(cudaMemcpy(sum_cov, ctx->g.path_cov_sum + sp.hap_base, sp.n_paths * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (sum_len) CU(cudaMemcpy(sum_len, ctx->g.path_len_sum + sp.hap_base, sp.n_paths * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return PTX_OK;
}

int ptx_hap_trio_counts(ptx_ctx* ctx, int s, int64_t* U, int64_t* nz) {
    int rc = check_results(ctx, s);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    const SpeciesHost& sp = ctx->sp[s];
    if (sp.n_paths == 0) return PTX_OK;
    if (U) {
        std::vector<uint64_t> ts((size_t)sp.n_paths + 1, 0);
        if (ctx->g.T > 0) CU(cudaMemcpy(ts.data(), ctx->g.trio_start + sp.hap_base, (sp.n_paths + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        for (int64_t h = 0; h < sp.n_paths; ++h) U[h] = (int64_t)(ts[h + 1] - ts[h]);
    }
    if (nz) CU(cudaMemcpy(nz, ctx->g.hap_nz + sp.hap_base, sp.n_paths * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return PTX_OK;
}

// gaf_filter.rs:44-97 on the GPU: newline index -> one thread per line (column split, i32/f64 parse) -> best
// (matches, identity) per read id in a 128-bit-CAS hash table -> first qualifying line per id -> compaction.
int ptx_filter_gaf(ptx_ctx* ctx, const uint8_t* bytes, size_t n, uint64_t* out_line_off, int64_t cap, int64_t* n_out) {
    if (!ctx || (!bytes && n) || !n_out || (cap > 0 && !out_line_off)) return fail(ctx, PTX_E_INVALID, "ptx_filter_gaf: bad arguments");
    cudaSetDevice(ctx->device);
    *n_out = 0;
    if (n == 0) return PTX_OK;
    cudaStream_t st = ctx->st;
    const bool add_nl = bytes[n - 1] != '\n';
    const uint64_t nn = n + (add_nl ? 1 : 0);
    const uint32_t n_micro = (uint32_t)((nn + MICRO - 1) / MICRO);
    uint8_t* d_text = nullptr;
    uint32_t* d_cnt = nullptr;
    uint64_t *d_base = nullptr, *d_scratch = nullptr, *d_line = nullptr, *d_scan = nullptr, *d_out = nullptr;
    uint32_t *d_sel = nullptr, *d_flags = nullptr;
    uint8_t* d_qual = nullptr;
    uint8_t* d_recs = nullptr;
    ulonglong2 *d_keys = nullptr, *d_best = nullptr;
    unsigned long long* d_first = nullptr;
    int rc = PTX_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        cudaFree(d_text); cudaFree(d_cnt); cudaFree(d_base); cudaFree(d_scratch); cudaFree(d_line); cudaFree(d_scan); cudaFree(d_out);
        cudaFree(d_sel); cudaFree(d_flags); cudaFree(d_qual); cudaFree(d_recs); cudaFree(d_keys); cudaFree(d_best); cudaFree(d_first);
    };
#define FLT(call) do { if ((rc = (call)) != PTX_OK) { cleanup(); return rc; } } while (0)
    FLT(dalloc(ctx, &d_text, (size_t)n_micro * MICRO + 64, false));
    if (cudaMemcpyAsync(d_text, bytes, n, cudaMemcpyHostToDevice, st) != cudaSuccess) { cleanup(); return fail(ctx, PTX_E_CUDA, "H2D copy failed"); }
    cudaMemsetAsync(d_text + n, '\n', (size_t)n_micro * MICRO + 64 - n, st);
    FLT(dalloc(ctx, &d_cnt, n_micro, false));
    FLT(dalloc(ctx, &d_base, (size_t)n_micro + 1, false));
    FLT(dalloc(ctx, &d_scratch, (size_t)n_micro / 2048 + 4, false));
    launch_flt_count_nl(d_text, nn, n_micro, d_cnt, st);
    launch_scan_u32(d_cnt, d_base, n_micro, d_scratch, st);
    uint64_t n_lines = 0;
    cudaMemcpyAsync(&n_lines, d_base + n_micro, sizeof n_lines, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); return fail(ctx, PTX_E_CUDA, "filter: newline index failed: %s", cudaGetErrorString(cudaGetLastError())); }
    FLT(dalloc(ctx, &d_line, (size_t)n_lines + 2, false));
    launch_flt_line_starts(d_text, nn, n_micro, d_base, d_line, st);
    const uint64_t tcap = 1ull << std::max<uint32_t>(10, log2_ceil(n_lines * 2 + 1));
    FLT(dalloc(ctx, &d_recs, (size_t)n_lines * flt_rec_bytes() + 64, false));
    FLT(dalloc(ctx, &d_keys, tcap));
    FLT(dalloc(ctx, &d_best, tcap));
    FLT(dalloc(ctx, &d_first, tcap, false));
    cudaMemsetAsync(d_first, 0xFF, tcap * sizeof(unsigned long long), st);
    FLT(dalloc(ctx, &d_qual, (size_t)n_lines + 1, false));
    FLT(dalloc(ctx, &d_sel, (size_t)n_lines + 1, false));
    FLT(dalloc(ctx, &d_flags, 4));
    FLT(dalloc(ctx, &d_scan, (size_t)n_lines + 2, false));
    launch_flt_pipeline(d_text, nn, d_line, n_lines, d_recs, d_keys, d_best, d_first, tcap - 1, 64 - log2_ceil(tcap), d_qual, d_sel, d_flags, st);
    dfree(d_scratch);
    FLT(dalloc(ctx, &d_scratch, (size_t)n_lines / 2048 + 4, false));
    launch_scan_u32(d_sel, d_scan, n_lines, d_scratch, st);
    uint64_t n_sel = 0;
    uint32_t h_flags[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(&n_sel, d_scan + n_lines, sizeof n_sel, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_flags, d_flags, sizeof h_flags, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); return fail(ctx, PTX_E_CUDA, "filter kernels failed: %s", cudaGetErrorString(cudaGetLastError())); }
    if (h_flags[2]) { cleanup(); return fail(ctx, PTX_E_UNSUPPORTED, "identity field outside the exactly-representable range (more than 15 significant digits, |exp10| > 22, inf or nan)"); }
    *n_out = (int64_t)n_sel;
    const uint64_t ncopy = std::min<uint64_t>(n_sel, cap > 0 ? (uint64_t)cap : 0);
    if (ncopy) {
        FLT(dalloc(ctx, &d_out, (size_t)ncopy, false));
        launch_flt_compact(d_sel, d_scan, d_line, n_lines, d_out, ncopy, st);
        cudaMemcpyAsync(out_line_off, d_out, ncopy * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    }
#undef FLT
    cleanup();
    return PTX_OK;
}

int ptx_comm_unique_id(void* out128) {
    if (!out128) return PTX_E_INVALID;
    if (!g_nccl.load()) return PTX_E_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return PTX_E_NCCL;
    memcpy(out128, &id, sizeof id);
    return PTX_OK;
}

int ptx_comm_init(ptx_ctx* ctx, int n_ranks, int rank, const void* id128) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, PTX_E_INVALID, "ptx_comm_init: bad arguments");
    if (!g_nccl.load()) return fail(ctx, PTX_E_NCCL, "libnccl.so.2 not found");
    if (ctx->total_records + ctx->pending_records > 0) return fail(ctx, PTX_E_STATE, "ptx_comm_init must precede the first ptx_ingest_gaf");
    cudaSetDevice(ctx->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    int rc = nccl_check(ctx, g_nccl.CommInitRank(&ctx->comm, n_ranks, id, rank), "ncclCommInitRank");
    if (rc) return rc;
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    // a second communicator over the same ranks carries the id-box exchange on the side stream, so that it really
    // runs beside the reductions of the main stream (operations on ONE communicator execute in issue order)
    ctx->comm_x = ctx->comm;
    if (g_nccl.CommSplit && !getenv("PTX_NO_COMM_SPLIT")) {
        ncclComm_t c2 = nullptr;
        if (g_nccl.CommSplit(ctx->comm, 0, rank, &c2, nullptr) == 0 && c2) ctx->comm_x = c2;
    }
    if (n_ranks > 1 && !getenv("PTX_NO_P2P")) {
        rc = p2p_setup(ctx);
        if (rc) return rc;
    }
    return PTX_OK;
}

// ---- several GPUs driven from ONE process (the reference CLI is one process) ---------------------------------------
// One context per device, NCCL communicators from ncclCommInitAll (two sets: reductions and the id-box exchange), and
// the id boxes in peer memory through cudaDeviceEnablePeerAccess - the in-process twin of ptx_comm_init + CUDA IPC.
int ptx_create_multi(const int* devices, int n_devices, int64_t expected_records_per_device, ptx_ctx** out) {
    if (!devices || !out || n_devices < 1) return PTX_E_INVALID;
    for (int i = 0; i < n_devices; ++i) out[i] = nullptr;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return PTX_E_INVALID;
    auto destroy_all = [&]() {
        for (int i = 0; i < n_devices; ++i) { if (out[i]) ptx_destroy(out[i]); out[i] = nullptr; }
    };
    for (int i = 0; i < n_devices; ++i) {
        const int rc = ptx_create(devices[i], &out[i]);
        if (rc) { destroy_all(); return rc; }
        out[i]->reserve_records = std::max<int64_t>(expected_records_per_device, 0);
    }
    if (n_devices == 1) return PTX_OK;
    if (!g_nccl.load() || !g_nccl.CommInitAll) { destroy_all(); return PTX_E_NCCL; }
    std::vector<ncclComm_t> c1(n_devices, nullptr), c2(n_devices, nullptr);
    if (g_nccl.CommInitAll(c1.data(), n_devices, devices) != 0 || g_nccl.CommInitAll(c2.data(), n_devices, devices) != 0) {
        destroy_all();
        return PTX_E_NCCL;
    }
    for (int i = 0; i < n_devices; ++i) {
        out[i]->comm = c1[i];
        out[i]->comm_x = c2[i];
        out[i]->n_ranks = n_devices;
        out[i]->rank = i;
    }
    // peer-memory id boxes: every device must reach every other one, and a size hint is needed (as in p2p_setup)
    bool peer_ok = expected_records_per_device > 0 && !getenv("PTX_NO_P2P");
    for (int i = 0; peer_ok && i < n_devices; ++i)
        for (int j = 0; peer_ok && j < n_devices; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) peer_ok = false;
        }
    for (int i = 0; i < n_devices; ++i) {
        ptx_ctx* ctx = out[i];
        cudaSetDevice(ctx->device);
        const int rc = xchg_ensure(ctx, 0);  // side stream, events, cursors, box pointer array
        if (rc) { destroy_all(); return rc; }
    }
    if (peer_ok) {
        const uint64_t P = (uint64_t)n_devices, hint = (uint64_t)expected_records_per_device;
        int64_t test_cap = 0;
        for (int i = 0; i < n_devices; ++i) { test_cap = std::max(test_cap, out[i]->test_box_cap); out[i]->test_box_cap = 0; }
        const uint64_t cap = test_cap > 0 ? (uint64_t)test_cap : hint / P + hint / (4 * P) + 4096;
        for (int i = 0; peer_ok && i < n_devices; ++i) {
            cudaSetDevice(devices[i]);
            for (int j = 0; j < n_devices; ++j) {
                if (i == j) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
                cudaGetLastError();
            }
            if (peer_ok && cudaMalloc((void**)&out[i]->p2p_inbox, (size_t)P * cap * sizeof(ulonglong2)) != cudaSuccess) { cudaGetLastError(); peer_ok = false; }
        }
        if (peer_ok) {
            for (int i = 0; i < n_devices; ++i) {
                ptx_ctx* ctx = out[i];
                cudaSetDevice(ctx->device);
                std::vector<ulonglong2*> ptrs(P, nullptr);
                for (int q = 0; q < n_devices; ++q)
                    if (q != i) ptrs[q] = out[q]->p2p_inbox + (uint64_t)i * cap;  // my slice of rank q's inbox (unified addressing)
                if (cudaMemcpy(ctx->d_box_ptr, ptrs.data(), (size_t)P * sizeof(ulonglong2*), cudaMemcpyHostToDevice) != cudaSuccess) { destroy_all(); return PTX_E_CUDA; }
                ctx->p2p = true;
                ctx->p2p_inprocess = true;
                ctx->p2p_cap = cap;
                ctx->box_cap = cap;
            }
        } else {
            for (int i = 0; i < n_devices; ++i) dfree(out[i]->p2p_inbox);
        }
    }
    return PTX_OK;
}

// ptx_finalize of every context of a ptx_create_multi group.  The collectives inside need all ranks at once: each
// context's finalize runs on its own host thread here; returns the first error code.
int ptx_finalize_multi(ptx_ctx* const* ctxs, int n) {
    if (!ctxs || n < 1) return PTX_E_INVALID;
    std::vector<int> rc((size_t)n, PTX_OK);
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back([&, i]() { rc[(size_t)i] = ptx_finalize(ctxs[i]); });
    rc[0] = ptx_finalize(ctxs[0]);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[(size_t)i]) return rc[(size_t)i];
    return PTX_OK;
}

int ptx_timing(ptx_ctx* ctx, double* ingest_ms, double* finalize_ms, int64_t* kernel_launches) {
    if (!ctx) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->st));
    if (ingest_ms) *ingest_ms = ev_sum(ctx->ev_count) + ev_sum(ctx->ev_ingest) + ev_sum(ctx->ev_apply);
    if (finalize_ms) *finalize_ms = ev_sum(ctx->ev_final);
    if (kernel_launches) *kernel_launches = kernel_launch_count();
    return PTX_OK;
}

int ptx_stats_json(ptx_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return PTX_E_INVALID;
    cudaSetDevice(ctx->device);
    chunks_resolve(ctx);
    cudaStreamSynchronize(ctx->st);
    size_t text = 0;
    size_t text_resident = 0;
    for (auto& ch : ctx->chunks) { text += ch.n; if (ch.buf) text_resident += ch.n; }
    snprintf(buf, cap,
             "{\"records\": %lld, \"chunks\": %zu, \"text_bytes\": %zu, \"nodes\": %lld, \"paths\": %lld, \"path_steps\": %lld, "
             "\"unique_trios\": %lld, \"bit_words\": %llu, \"id_set_slots\": %llu, \"ids_unique\": %d, \"mixed_groups\": %d, "
             "\"count_ms\": %.4f, \"ingest_ms\": %.4f, \"apply_ms\": %.4f, \"ingest_launches\": %zu, \"finalize_ms\": %.4f, \"kernel_launches\": %lld, \"ranks\": %d, \"p2p_boxes\": %d, \"table_allocs\": %lld, \"text_buffers_released\": %zu, \"text_bytes_resident\": %zu, \"graph_commit_ms\": %.3f}",
             (long long)ctx->total_records, ctx->chunks.size(), text, (long long)ctx->g.N, (long long)ctx->g.Htot, (long long)ctx->g.P,
             (long long)ctx->g.T, (unsigned long long)ctx->g.n_bit_words, (unsigned long long)ctx->ds_cap, ctx->h_flags[0] == 0 ? 1 : 0,
             ctx->h_flags[1] != 0 ? 1 : 0, ev_sum(ctx->ev_count), ev_sum(ctx->ev_ingest), ev_sum(ctx->ev_apply), ctx->ev_ingest.size(), ev_sum(ctx->ev_final), (long long)kernel_launch_count(), ctx->n_ranks, ctx->p2p ? 1 : 0, (long long)ctx->n_table_allocs, ctx->text_pool.size(), text_resident, ctx->commit_ms);
    return PTX_OK;
}

}  // extern "C"
