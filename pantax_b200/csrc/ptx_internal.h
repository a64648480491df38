// Internal structs shared by ptx_kernels.cu (device code + launchers) and ptx_api.cu
// (context, C ABI).  Not installed; the public surface is include/pantax_gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx_core.cuh"

namespace ptx {

// ---- tile geometry of the GAF scan -------------------------------------------------
// A chunk buffer is  [PRE '\n' bytes][text, whole lines][ '\n' padding ].
// The text is cut into 4 KB micro-tiles; K1 counts the records of each (a record belongs to the
// micro-tile that holds the newline BEFORE it; the first record of the chunk, preceded by the PRE
// padding, belongs to micro-tile 0).  An ingest tile is a run of `rows_per_warp` (1..8) micro-tiles
// (each of the CTA's 8 warps scans rows_per_warp rows of 512 B), chosen per chunk from the measured
// mean line length so that a tile holds about one record per thread.  A record therefore starts in
// (t*tile, (t+1)*tile]; the tile is staged in shared memory as text[t*tile, t*tile + tile + OVER).
constexpr uint32_t MICRO = 4096;
constexpr uint32_t MAX_ROWS = 8;  // rows of 512 B a warp scans per tile
constexpr uint32_t MAX_TILE = MAX_ROWS * MICRO;
constexpr uint32_t OVER = 2048;
constexpr uint32_t PRE = 256;  // keeps `text` 256-byte aligned inside a cudaMalloc'ed buffer
constexpr int INGEST_THREADS = 256;
#ifndef PTX_SHORT_THREADS
#define PTX_SHORT_THREADS 128
#endif
constexpr int SHORT_THREADS = PTX_SHORT_THREADS;       // k_ingest_s: small CTAs, 7 per SM - the barriers between its phases overlap across CTAs
#ifndef PTX_APPLY_MINB
#define PTX_APPLY_MINB 6  // k_apply: resident CTAs of 256 threads per SM the register allocation aims at: 6 -> 40 registers, no spills (0.923 ms; 8 -> 32 registers with a 48-byte stack: 0.931 ms; 4 -> 46 registers: 1.019 ms)
#endif
#ifndef PTX_SHORT_RECCAP
#define PTX_SHORT_RECCAP 512
#endif
constexpr uint32_t SHORT_REC_CAP = PTX_SHORT_RECCAP;  // k_ingest_s: line starts kept in smem per round
#ifndef PTX_LONG_MINB
#define PTX_LONG_MINB 4
#endif
constexpr uint32_t LONG_REC_CAP = 512;   // k_ingest_l: line starts kept in smem per round (4 CTAs of 55 KB per SM)
#ifndef PTX_LONG_THREADS
#define PTX_LONG_THREADS 256
#endif
#ifndef PTX_LONG_TILE_MAX
#define PTX_LONG_TILE_MAX (7 * 4096)
#endif
constexpr int LONG_THREADS = PTX_LONG_THREADS;  // k_ingest_l
constexpr uint32_t LONG_TILE_MAX = PTX_LONG_TILE_MAX;  // k_ingest_l: largest tile (4 resident CTAs per SM)
constexpr uint32_t REC_CAP = 1024;  // record starts kept in smem per round
constexpr double LONG_LINE_BYTES = 320.0;  // mean line length from which a chunk is parsed by k_ingest<LONG>
constexpr uint32_t HIST_SLOTS_LOG2 = 7;  // per-tile species-count accumulators in shared memory (multi-species runs)
constexpr uint32_t HIST_SLOTS = 1u << HIST_SLOTS_LOG2;
constexpr uint64_t BITS_SLICE_PAD = 4 * 64;  // spare bitmap words: n_bit_words rounded up to n_ranks slices of a multiple of 4 words (<= 64 ranks)
#ifndef PTX_SHORT_STASH
#define PTX_SHORT_STASH 16
#endif
#ifndef PTX_SHORT_MINB
#define PTX_SHORT_MINB 7
#endif
constexpr uint32_t SHORT_STASH_CAP = PTX_SHORT_STASH;  // k_ingest_s: walk nodes per record kept in smem (a longer walk is decoded again at write-out).  7 CTAs per SM: measured faster than 8 (1.05 vs 1.34 ms) and than 6 (1.12 ms)
constexpr uint32_t STASH_CAP = 16;  // walk nodes per record kept in smem between the parse and the coverage pass

// record-table flags (IngestArgs::meta_b[e].y): walk length in the low 24 bits
constexpr uint32_t RM_W_MASK = 0x00FFFFFFu;
constexpr uint32_t RM_LABELLED = 0x80000000u;  // species label != U
constexpr uint32_t RM_ELIGIBLE = 0x40000000u;  // path, c7, c8, c9 all non-null (profile.rs:380-399)
constexpr uint32_t RM_MONOTONE = 0x20000000u;  // strictly monotone node ids: no node repeats in the walk
constexpr uint32_t RM_VALID = 0x10000000u;     // the line slot is a GAF row (not an empty line / '@' comment)

// mode flags of k_apply
constexpr int MODE_CLASSIFY = 1;  // labels, species counts, read-id set insert
constexpr int MODE_COVER = 2;     // node coverage / trio accumulation
constexpr int MODE_KEEPMASK = 4;  // skip reads whose id group is DS_MIXED (replay pass)
constexpr uint32_t BOX_STAGE_RANKS = 16;  // k_apply groups box entries per destination in shared memory up to this many ranks
constexpr int MODE_REBOX = 8;     // multi-GPU, rare: fill the outboxes again after they were enlarged (no id-set insert)

struct GraphDev {
    // nodes (all uploaded species concatenated; g = node_base[s] + local id)
    int64_t N = 0;
    uint32_t* len = nullptr;          // [N]
    uint64_t* bit_off = nullptr;      // [N+1] exclusive prefix of len
    unsigned long long* bases = nullptr;  // [N] int64 accumulators
    uint4* ninfo = nullptr;           // [N] {len, flags (NI_FULL | NI_TRIO_MID), bit_off lo, hi}: the one gather of the coverage pass
    uint8_t* full = nullptr;          // [N] NI_FULL extracted at finalize (staging for k_cov and the cross-rank max)
    uint32_t* bits = nullptr;         // [ceil(total_bits/32)+1] packed per-base covered bitmap
    uint64_t n_bit_words = 0;         // words in use; the allocation has BITS_SLICE_PAD more (the per-rank slices of the OR-reduction are rounded up)
    uint32_t* cov = nullptr;          // [N] covered bases (finalize)
    // paths
    int64_t Htot = 0, P = 0;          // paths, total steps
    uint32_t* pnode = nullptr;        // [P] global node idx, bit31 = node already seen earlier in this path
    uint64_t* poff = nullptr;         // [Htot+1]
    unsigned long long* path_len_sum = nullptr;  // [Htot] sum len over distinct nodes
    unsigned long long* path_cov_sum = nullptr;  // [Htot] sum cov over distinct nodes (finalize)
    // unique trios
    int64_t T = 0;
    uint32_t* trio_key = nullptr;     // [T*3] canonical global node idx
    int64_t* trio_len = nullptr;      // [T]
    uint32_t* trio_owner = nullptr;   // [T] global hap idx
    unsigned long long* trio_bases = nullptr;  // [T]
    uint64_t* trio_start = nullptr;   // [Htot+1] first trio idx of each hap
    unsigned long long* hap_nz = nullptr;  // [Htot]
    uint4* tt = nullptr;              // probe table {a,b,c,idx}, a == TT_EMPTY: empty
    uint32_t tt_mask = 0;
};

struct IngestArgs {
    const uint8_t* text;   // chunk text (buffer + PRE)
    uint64_t n_bytes;      // text bytes (whole lines)
    uint64_t padded_bytes; // bytes readable from `text` (text + newline padding)
    uint32_t n_tiles;
    uint32_t tile_bytes;         // bytes of text per CTA: a multiple of 512 (k_ingest_s) / of 4096 (k_ingest<>, and whenever micro_base is used)
    uint32_t over_bytes;         // k_ingest_s: bytes staged behind the tile (the tail of its last line): 512..OVER, multiple of 512
    uint32_t pf_dist;            // k_ingest_s / k_ingest_l: a CTA prefetches the text of tile blockIdx + pf_dist into the L2 (0: off)
    uint32_t no_sort;            // k_ingest_s: keep the lines of a tile in file order (PTX_NO_SORT=1, measurements)
    uint32_t long_mode;          // long lines: warp-cooperative walk decode (k_ingest<true>)
    uint32_t long_new;           // PTX_LONG_NEW=1: k_ingest_l instead of k_ingest<true> for long lines (measurements)
    uint32_t old_short;          // PTX_OLD_INGEST=1: the byte-at-a-time short-read kernel of round 1 (A/B measurements)
    const uint64_t* micro_base;  // [n_micro] exclusive record prefix per micro-tile within the chunk, from the count pass.
                                 // null = SINGLE-PASS mode: no count pass ran, rows are numbered later from tile_info/row_key
    uint4* tile_info;            // single-pass: [n_tiles] {first record-table entry, line slots, GAF rows, 0} of each tile
    uint16_t* row_key;           // single-pass: [slots_cap] index of the entry's GAF row within its tile
    uint32_t slots_cap;          // single-pass: capacity of the record table (estimated; overflow -> cursors[3])
    uint32_t* labels;           // [chunk records]
    const uint32_t* labels_in;  // [chunk records] caller-supplied species labels (ptx_ingest_labels), or null: classify
    // record table + CSR walks written by k_ingest, consumed by k_apply (entries in length-sorted tile order)
    uint4* meta_b;              // [line slots] {node_off, flags|W, label, id hash hi}
    longlong2* meta_a;          // [line slots] {c8, c9} of eligible records
    unsigned long long* hash_lo;  // [line slots] id hash lo of labelled records
    uint32_t* nodes;            // [<= text bytes / 2] raw node ids of eligible records' walks
    uint32_t nodes_cap;         // capacity of `nodes` (k_ingest_s abandons the chunk instead of writing beyond it)
    uint32_t* cursors;          // [0] next record-table entry, [1] next node slot, [2] GAF rows, [3] single-pass estimate too small, [4] labelled rows
    // multi-GPU: id entries routed to the rank owning their hash (written by k_apply<CLASSIFY>, sent at finalize)
    // multi-GPU (null on one GPU): box_ptr[q] = where {hash, state} of records whose id rank q owns are appended -
    // this rank's slice of q's inbox in PEER memory (stores travel over NVLink while k_apply runs), or a local outbox
    // that ptx_finalize sends with ncclSend when peer memory is not set up
    ulonglong2* const* box_ptr;
    unsigned long long* out_cursor;  // [n_ranks]
    uint64_t box_cap;
    uint32_t n_ranks, rank;
    RangesView ranges;
    unsigned long long* hist;   // [hist_copies][S*4]: tile t adds to copy t % hist_copies (single-pass chunks; 1 copy otherwise)
    uint32_t hist_copies, hist_stride;
    ulonglong2* ds;             // read-id set slots
    uint32_t ds_shift;          // 64 - log2(capacity)
    uint64_t ds_mask;
    // node-coverage scatter variant of k_apply<COVER> (north_star stage 2; profiles/r2_scatter_bakeoff.md):
    //   0 one RED.ADD.64 per lane and node   1 lanes of a warp that hit the same node add once (match.any)
    //   2 per-CTA shared-memory table absorbing repeated nodes, flushed once per CTA
    //   3 (node, bases) pairs written beside the CSR walk, then radix sort by node + segmented reduce (launch_scatter_sorted)
    uint32_t scatter_var;
    uint32_t* pair_key;             // variant 3: [node slots] node index per CSR slot (0xFFFFFFFF: nothing to add)
    unsigned long long* pair_val;   // variant 3: [node slots]
    uint32_t* flags;            // [0] dup id seen, [1] mixed-species id group seen, [2] exchange box overflow, [3] supplied label outside its range
    uint32_t* err;              // [S] bit0: profile.rs:854 tripped
    // coverage
    uint4* ninfo;
    unsigned long long* bases;
    uint32_t* bits;
    const uint4* tt;
    uint32_t tt_mask;
    unsigned long long* trio_bases;
    // L2 eviction policies (createpolicy encodings) of the two kinds of traffic: what is read once - GAF text, record table,
    // id-set slots - is fetched evict-first, the graph arrays every read gathers from (ninfo, bases, bitmap, trio table)
    // evict-last, so that 1.7 GB of streams per step do not push 30 MB of node arrays out of the L2
    uint64_t pol_keep, pol_stream, pol_ds;
    uint32_t ds_epoch;      // id-set slots of another epoch are empty (a new pass increments it instead of clearing the table)
};

// launchers (ptx_kernels.cu); all asynchronous on `st`
void launch_count_records(const uint8_t* text, uint64_t n_bytes, uint32_t n_micro, uint32_t* micro_count, unsigned long long* total_slots,
                          cudaStream_t st);
void launch_ingest(const IngestArgs& a, cudaStream_t st);
constexpr uint32_t ENTRIES_FROM_DEVICE = 0xFFFFFFFFu;  // launch_apply: take the entry count (and the abandon flag) from a.cursors
void launch_apply(const IngestArgs& a, uint32_t n_entries, int mode, cudaStream_t st);
void launch_count_labelled(const unsigned long long* hist, uint32_t S, unsigned long long* dst, cudaStream_t st);
void launch_hist_merge(const unsigned long long* chunk_hist, unsigned long long* hist, uint32_t n, uint32_t copies, uint32_t* cursors, cudaStream_t st);
void launch_tile_rows(const uint4* tile_info, uint32_t* rows, uint32_t n_tiles, cudaStream_t st);
void launch_labels_from_table(const uint4* tile_info, const uint64_t* tile_off, const uint4* meta_b, const uint16_t* row_key, uint32_t* labels,
                              uint32_t n_tiles, cudaStream_t st);
void launch_ds_rehash(const ulonglong2* old_slots, uint64_t old_cap, ulonglong2* new_slots, uint32_t new_shift,
                      uint64_t new_mask, uint32_t ep, cudaStream_t st);
void launch_ninfo_build(const uint32_t* len, const uint64_t* bit_off, uint4* ninfo, int64_t N, cudaStream_t st);
void launch_ninfo_full(uint4* ninfo, uint8_t* full, int64_t N, int mode, cudaStream_t st);
// cross-rank id groups (multi-GPU finalize)
void launch_ds_collect_mixed(const ulonglong2* slots, uint64_t cap, uint32_t ep, unsigned long long* cursor, ulonglong2* out, uint64_t out_cap, cudaStream_t st);
void launch_ds_merge_boxes(const ulonglong2* inbox, const unsigned long long* off, const unsigned long long* cnt, uint32_t n_boxes, uint64_t max_cnt,
                           ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, uint32_t* flags, cudaStream_t st);
void launch_ds_apply_mixed(const ulonglong2* in, uint64_t n, ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, uint32_t* flags, cudaStream_t st);

// graph commit
void launch_mark_path_dups(uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint64_t* pbm_off, const uint32_t* pbase, uint32_t* bm,
                           cudaStream_t st);
void launch_path_len_sum(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint32_t* len,
                         unsigned long long* out, cudaStream_t st);
void launch_trio_count(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, uint4* keys, uint32_t* cnt,
                       uint32_t mask, cudaStream_t st);
void launch_trio_flag(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint4* keys,
                      const uint32_t* cnt, uint32_t mask, uint32_t* flag, cudaStream_t st);
// exclusive scan of n uint32 -> uint64 prefix (out[n] = total); scratch >= (n/2048+2) uint64
void launch_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch, cudaStream_t st);
void launch_trio_emit(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint32_t* flag,
                      const uint64_t* scan, const uint32_t* len, uint32_t* trio_key, int64_t* trio_len,
                      uint32_t* trio_owner, uint4* tt, uint32_t tt_mask, uint4* ninfo, uint64_t* trio_start,
                      cudaStream_t st);

// finalize
void launch_cov(const GraphDev& g, bool bits_only, cudaStream_t st);  // bits_only: the full-node flags were written into the bitmap (multi-GPU)
void launch_path_cov_sum(const GraphDev& g, cudaStream_t st);
void launch_hap_nz(const GraphDev& g, cudaStream_t st);
void launch_depth(const unsigned long long* num, const uint32_t* den32, const int64_t* den64, double* out, uint64_t n,
                  cudaStream_t st);
// OR-merge a peer's bitmap / full flags into ours (multi-GPU finalize)
void launch_or_words(uint32_t* dst, const uint32_t* src, uint64_t n_words, cudaStream_t st);
void launch_bits_fill_full(const GraphDev& g, cudaStream_t st);
void launch_or_slices(uint32_t* dst, const uint32_t* src, uint32_t n_src, uint64_t n_words, cudaStream_t st);

// GFA graph text on the device (section 8f3)
constexpr uint32_t GFA_PIECE_BYTES = 4096;
struct GfaPathLineHost { unsigned long long line, name_beg, fld_beg, fld_end; uint32_t name_len, kind, pad0, pad1; };
size_t gfa_path_line_bytes();
void launch_gfa_field_end(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, const uint32_t* piece_line, unsigned long long* line_tab,
                          uint32_t n_pieces, cudaStream_t st);
void launch_gfa_lines(const uint8_t* text, const uint64_t* line_off, uint64_t n_lines, int64_t n_nodes, uint32_t* len, uint32_t* is_s, uint32_t* s_adj,
                      void* plist, unsigned long long* pcount, uint64_t pcap, uint32_t* flags, cudaStream_t st);
void launch_gfa_check_order(const uint32_t* is_s, const uint32_t* s_adj, const uint64_t* ord, uint64_t n_lines, uint32_t* flags, cudaStream_t st);
void launch_gfa_path_count(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, uint32_t* piece_cnt, uint32_t n_pieces, uint32_t* flags,
                           cudaStream_t st);
void launch_gfa_path_decode(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, const uint64_t* piece_fbeg, const uint64_t* piece_dst,
                            int64_t n_nodes, uint32_t* out, uint32_t n_pieces, uint32_t* flags, cudaStream_t st);

// K10 long-read filter
void launch_flt_count_nl(const uint8_t* text, uint64_t n, uint32_t n_micro, uint32_t* cnt, cudaStream_t st);
void launch_flt_line_starts(const uint8_t* text, uint64_t n, uint32_t n_micro, const uint64_t* micro_base, uint64_t* line_off, cudaStream_t st);
size_t flt_rec_bytes();
void launch_flt_pipeline(const uint8_t* text, uint64_t n, const uint64_t* line_off, uint64_t n_lines, void* recs, ulonglong2* keys,
                         ulonglong2* best, unsigned long long* first, uint64_t mask, uint32_t shift, uint8_t* qual, uint32_t* sel,
                         uint32_t* flags, cudaStream_t st);
void launch_flt_compact(const uint32_t* sel, const uint64_t* scan, const uint64_t* line_off, uint64_t n_lines, uint64_t* out, uint64_t cap,
                        cudaStream_t st);

// scatter variant 3: sort the n (node, bases) pairs of a chunk by node and add every run's sum to bases[] (CUB radix sort +
// reduce-by-key, then one kernel).  `tmp` holds tmp_bytes of scratch (scatter_sorted_tmp_bytes(n)).
size_t scatter_sorted_tmp_bytes(uint64_t n);
void launch_scatter_sorted(const uint32_t* key, const unsigned long long* val, uint64_t n, unsigned long long* bases, void* tmp, size_t tmp_bytes,
                           cudaStream_t st);

int64_t kernel_launch_count();

}  // namespace ptx
