/* pantax_gpu.h - C ABI of the B200-native PanTax alignment-to-abundance hot path.
 *
 * The reference (LuoGroup2023/PanTax v2.1.0) has no FFI seam around this path: every
 * hot function is a private Rust fn.  This header DEFINES the drop-in boundary; each
 * entry point names the reference function(s) it replaces (paths relative to
 * pantax/src/).  INTEGRATION.md shows the `extern "C"` block + call-site patch a
 * maintainer adds to profile.rs, and rust/pantax-gpu-sys/ holds that shim as source.
 *
 * Model: one ptx_ctx per process per GPU (one process per GPU; multi-GPU runs shard
 * the GAF by read batch and reduce inside ptx_finalize once ptx_comm_init was called).
 * A ctx is driven by ONE host thread.  All functions return 0 (PTX_OK) or a negative
 * PTX_E_* code; nothing aborts or throws across the boundary; ptx_last_error() gives
 * a NUL-terminated message owned by the ctx.  There is no CPU fallback: without a
 * CUDA device ptx_create fails with PTX_E_CUDA.
 *
 * Call order:
 *   ptx_create -> ptx_set_ranges -> [ptx_upload_graph x species -> ptx_commit_graphs]
 *   -> ptx_ingest_gaf* (any number of chunks) -> ptx_finalize -> getters.
 * Graphs may also be uploaded AFTER the ingest (the reference's own order: species
 * abundance first, then only the abundant species' graphs, profile.rs:3359-3363):
 *   ... ptx_ingest_gaf* -> ptx_finalize -> ptx_species_counts -> ptx_upload_graph x k
 *   -> ptx_commit_graphs -> ptx_finalize (coverage pass over the retained record tables).
 */
#ifndef PANTAX_GPU_H
#define PANTAX_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptx_ctx ptx_ctx;

#define PTX_OK 0
#define PTX_E_INVALID (-1)        /* bad argument */
#define PTX_E_CUDA (-2)           /* CUDA runtime / no device */
#define PTX_E_NOMEM (-3)          /* host or device allocation failed */
#define PTX_E_STATE (-4)          /* call order violated */
#define PTX_E_NODE_ORDER (-5)     /* profile.rs:489 / zip.rs:103  node ids not consecutive */
#define PTX_E_ZERO_LEN (-6)       /* profile.rs:494 / :985        node length 0 */
#define PTX_E_START_GT_LEN (-7)   /* profile.rs:854               read start beyond first node (reference panics) */
#define PTX_E_NVERT_MISMATCH (-8) /* profile.rs:2938              range size != number of nodes */
#define PTX_E_NO_GRAPH (-9)       /* getter for a species without an uploaded graph */
#define PTX_E_NCCL (-10)
#define PTX_E_IO (-11)
#define PTX_E_RANGE (-12)         /* species / buffer index out of range */
#define PTX_E_UNSUPPORTED (-13)

#define PTX_LABEL_UNCLASSIFIED 0xFFFFFFFFu /* the reference's "U" (rcls.rs:257) */

/* ---- lifetime --------------------------------------------------------------------- */
int ptx_create(int device, ptx_ctx** out);
void ptx_destroy(ptx_ctx* ctx);
const char* ptx_last_error(const ptx_ctx* ctx);
const char* ptx_version(void);

/* ---- inputs ----------------------------------------------------------------------- */
/* species_range.txt rows in FILE ORDER: taxid, start, end (1-based inclusive global node
 * ids).  Replaces rcls::load_species_range (rcls.rs:40-71).  Strings are copied. */
int ptx_set_ranges(ptx_ctx* ctx, int n_species, const char* const* taxid, const int64_t* start, const int64_t* end);

/* One species' Graph (types.rs:51-55) as loaded by zip::load_from_zip_graph
 * (zip.rs:236-262) or profile::read_gfa (profile.rs:466-545): nodes_len[n] (local id =
 * GFA id - 1), H paths in hap-NAME order (BTreeMap), path h = path_nodes[path_off[h] ..
 * path_off[h+1]) local ids.  Caller keeps ownership; data are copied.
 * Errors: PTX_E_NVERT_MISMATCH, PTX_E_ZERO_LEN, PTX_E_INVALID (node id >= n). */
int ptx_upload_graph(ptx_ctx* ctx, int species, const int64_t* nodes_len, int64_t n, const uint64_t* path_off,
                     const uint64_t* path_nodes, int64_t n_paths);

/* The same graph straight from the species' GFA text (species_gfa/<taxid>.gfa), parsed ON THE DEVICE: replaces
 * profile::read_gfa (profile.rs:466-545, previous = 0) - S lines in id order give nodes_len (sequence length), the digit runs of
 * the P (`\d+` of field 3, haplotype = field 2 up to '#') and W (`\d+` of the last field, haplotype = field 2) lines give the
 * paths; lines of one haplotype are concatenated in file order, haplotypes in name order.  `gfa` is host memory (n bytes).
 * Errors: PTX_E_NODE_ORDER (profile.rs:489), PTX_E_ZERO_LEN (:494), PTX_E_NVERT_MISMATCH, PTX_E_INVALID. */
int ptx_upload_graph_gfa(ptx_ctx* ctx, int species, const uint8_t* gfa, size_t n);

/* The Graph (types.rs:51-55) the library holds for a species after either upload - what profile::read_gfa returns: nodes_len[n],
 * path_off[n_paths + 1], path_nodes[ptx_species_path_steps] (local ids); null pointers are skipped.  ptx_species_path_name: the
 * haplotype id of path h (BTreeMap key; only known for ptx_upload_graph_gfa), returns its length. */
int ptx_species_graph(ptx_ctx* ctx, int species, int64_t* nodes_len, uint64_t* path_off, uint64_t* path_nodes);
int64_t ptx_species_path_steps(const ptx_ctx* ctx, int species);
int ptx_species_path_name(ptx_ctx* ctx, int species, int64_t h, char* buf, size_t cap);

/* Builds the device graph: node arrays, path CSR with distinct-node marks, and the
 * unique-trio table.  Replaces profile::trio_nodes_info (profile.rs:658-740) for all
 * uploaded species at once. */
int ptx_commit_graphs(ptx_ctx* ctx);

/* Optional hint: expected number of GAF records of this ctx (sizes the read-id set once).  Given on every
 * rank BEFORE ptx_comm_init it also sizes the peer-memory id boxes of the multi-GPU exchange (below). */
int ptx_reserve(ptx_ctx* ctx, int64_t expected_records);

/* Pinned host memory for the caller's GAF buffer (full-speed H2D).  The result getters below accept any host
 * pointer; one that comes from ptx_host_alloc makes their device->host copy a direct DMA as well. */
int ptx_host_alloc(size_t bytes, void** out);
int ptx_host_free(void* p);

/* GAF text from HOST memory.  Chunks may split lines anywhere; the library carries the
 * partial last line to the next call.  is_last != 0 on the final chunk.  Replaces
 * rcls::load_gaf_file_lazy + process_reads_parallel_simple (rcls.rs:119-146, 306-323),
 * the integer part of species_profiling (profile.rs:208-297), group_reads_by_species
 * (profile.rs:361-463) and the read loop of get_node_abundances (profile.rs:787-919).
 * Asynchronous: returns when `bytes` has been copied (the caller may reuse the buffer), not when the
 * kernels are done.  After the first chunk of a ctx the text is read in a single pass and the record
 * counts come back later; ptx_finalize, ptx_num_records and the getters wait for them. */
int ptx_ingest_gaf(ptx_ctx* ctx, const uint8_t* bytes, size_t n, int is_last);

/* Strain-only resume (profile.rs:3365-3419: `--strain` without `--species`): the species column is read
 * from reads_classification.tsv (column 3, row-aligned with the GAF, profile.rs:3367-3385) instead of
 * being derived from the walk.  labels[i] = species index (order of ptx_set_ranges) or PTX_LABEL_UNCLASSIFIED for
 * GAF row i of this ctx's input, counted over all ptx_ingest_gaf calls; successive calls append.  Must
 * precede the ptx_ingest_gaf call that carries the row.  Cleared by ptx_reset and ptx_set_ranges. */
int ptx_ingest_labels(ptx_ctx* ctx, const uint32_t* labels, int64_t n);

/* GAF text already in DEVICE memory: obtain a padded device buffer, fill it (whole lines
 * only, e.g. cudaMemcpy or a device-side generator) and hand it over zero-copy.  The
 * buffer stays owned by the ctx and must not be modified until ptx_reset/ptx_destroy. */
int ptx_gaf_buffer_alloc(ptx_ctx* ctx, size_t capacity, int* buffer_id, void** device_ptr);
int ptx_ingest_gaf_device(ptx_ctx* ctx, int buffer_id, size_t n);

/* Runs what is still pending: cross-GPU reductions (if ptx_comm_init was called), the
 * duplicate-id rule (profile.rs:406-437; replays the coverage from the record table kept on the
 * device - no text is re-read - only if a mixed-species id group exists), covered-base counts (profile.rs:1018-1023), per-path sums
 * (profile.rs:2705-2729) and per-hap unique-trio counts (profile.rs:1112-1135). */
int ptx_finalize(ptx_ctx* ctx);

/* Forget all ingested records and accumulators; ranges and graphs stay. */
int ptx_reset(ptx_ctx* ctx);
/* Same, but the device GAF buffers (ptx_gaf_buffer_alloc) and their text are kept and may be
 * ingested again with ptx_ingest_gaf_device: another pass over text already resident in HBM. */
int ptx_rewind(ptx_ctx* ctx);

/* ---- outputs (caller-allocated buffers) -------------------------------------------- */
int64_t ptx_num_records(const ptx_ctx* ctx);   /* GAF rows (non-comment, non-empty lines), this rank; waits for in-flight chunks */
int ptx_num_species(const ptx_ctx* ctx);
int ptx_ids_unique(const ptx_ctx* ctx);        /* profile.rs:376 `unique` over all non-U rows */

/* labels[num_records]: species index (row of ptx_set_ranges) or PTX_LABEL_UNCLASSIFIED,
 * in GAF row order - the "species" column of rcls_profile's DataFrame (rcls.rs:320). */
int ptx_read_labels(ptx_ctx* ctx, uint32_t* labels);

/* counts[n_species][4] = read_count, sum(read_len), #(3<=mapq<=60), #(mapq==60)
 * (profile.rs:219-232, 264-277). */
int ptx_species_counts(ptx_ctx* ctx, int64_t* counts);

/* profile.rs:311-322: are the read lengths of the first 1000 non-U rows all equal? */
int ptx_equal_length(ptx_ctx* ctx, int* is_equal, int64_t* read_len);

int64_t ptx_species_nodes(const ptx_ctx* ctx, int species);  /* n, or <0 */
int64_t ptx_species_paths(const ptx_ctx* ctx, int species);  /* H, or <0 */
int64_t ptx_species_trios(const ptx_ctx* ctx, int species);  /* T, or <0 */

/* get_node_abundances (profile.rs:743-1026) outputs for one species.  Any of these
 * returns PTX_E_START_GT_LEN if a kept read of that species tripped profile.rs:854. */
int ptx_node_bases(ptx_ctx* ctx, int species, int64_t* bases);     /* bases_per_node[n]      :829,:881 */
int ptx_node_cov(ptx_ctx* ctx, int species, uint64_t* cov);        /* node_base_cov[n]       :1018-1023 */
int ptx_node_depth(ptx_ctx* ctx, int species, double* depth);      /* node_abundance_vec[n]  :980-990 */
int ptx_trio_bases(ptx_ctx* ctx, int species, int64_t* bases);     /* trio base counts[T]    :906 */
int ptx_trio_depth(ptx_ctx* ctx, int species, double* depth);      /* trio_node_abundance[T] :1006-1016 */

/* Unique trios of one species in the library's deterministic order (owner hap in name
 * order, window position): keys3[T][3] canonical local ids (min(a,c), b, max(a,c)),
 * len[T] = len[a]+len[b]+len[c], owner[T] = hap index.  Any pointer may be NULL. */
int ptx_trio_table(ptx_ctx* ctx, int species, uint64_t* keys3, int64_t* len, uint32_t* owner);

/* The REFERENCE's numbering of the same unique trios: order[i] = row of ptx_trio_table that profile.rs:705-716 numbers i, i.e. the
 * iteration order of the FxHashSet of profile.rs:659-685 (fxhash 0.2.1 over the three usize fields; the standard library's
 * hashbrown table, x86-64 group width 16; one `extend` per haplotype in name order) restricted to the trios that occur once.
 * Visible only in the f64 summation order of zscore_filter / frequencies_mean (profile.rs:1037-1041, 1146): a caller that wants the
 * reference's bits visits a haplotype's trio abundances in this order.  Host-only, no ctx: path_off[n_paths + 1] / path_nodes as given
 * to ptx_upload_graph, keys3[n_trios][3] as returned by ptx_trio_table.  PTX_E_INVALID when the two do not describe the same trios. */
int ptx_trio_ref_order(const uint64_t* path_off, const uint64_t* path_nodes, int64_t n_paths, const uint64_t* keys3, int64_t n_trios,
                       uint64_t* order);

/* Per path (name order): sum of node_base_cov and of nodes_len over the DISTINCT nodes
 * of the path - the two products of profile.rs:2714-2724 as exact integers. */
int ptx_path_sums(ptx_ctx* ctx, int species, int64_t* sum_cov, int64_t* sum_len);

/* Per hap: U = #unique trios owned, nz = # of those with abundance > 0 (profile.rs:1114-1135). */
int ptx_hap_trio_counts(ptx_ctx* ctx, int species, int64_t* U, int64_t* nz);

/* ---- long-read pre-filter ----------------------------------------------------------- */
/* gaf_filter::filter_max_alignment_mt (gaf_filter.rs:44-97) on a whole GAF in host memory:
 * writes the byte offsets of the kept lines (file order; ties -> first line in file order)
 * to out_line_off[cap] and their number to *n_out. */
int ptx_filter_gaf(ptx_ctx* ctx, const uint8_t* bytes, size_t n, uint64_t* out_line_off, int64_t cap, int64_t* n_out);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------ */
/* NCCL bootstrap: rank 0 calls ptx_comm_unique_id, broadcasts the 128 bytes by any means,
 * every rank calls ptx_comm_init.  ptx_finalize then all-reduces the int64 accumulators
 * and OR-reduces the covered-base bitmap; every rank ends with the global result.
 * Read-id groups (profile.rs:361-463) ignore shard boundaries: an id is kept only by the rank that owns its
 * hash.  If every rank called ptx_reserve first, the ranks map each other's inboxes through CUDA IPC and the
 * ingest kernels store foreign ids straight into the owner's memory over NVLink/NVSwitch (all ranks on one node);
 * otherwise, or if an inbox slice overflows, the ids travel with ncclSend/ncclRecv inside ptx_finalize. */
int ptx_comm_unique_id(void* out128);
int ptx_comm_init(ptx_ctx* ctx, int n_ranks, int rank, const void* id128);

/* Several GPUs from ONE process (the reference CLI, main.rs:32-58, is a single process): creates one context per
 * device of `devices[n_devices]` into out[n_devices], already joined into one communicator group (ncclCommInitAll;
 * rank = position in `devices`), with the id boxes in peer memory (cudaDeviceEnablePeerAccess) when
 * expected_records_per_device > 0 and every device can reach every other one.  Each context is then fed its own read
 * batch with ptx_ingest_gaf* (from one host thread or from one thread per context - a context is never shared between
 * threads), graphs and ranges are set on every context, and ptx_finalize_multi runs ptx_finalize of all of them (the
 * collectives inside need every rank at once; it uses one host thread per context).  Replaces the nested rayon loops
 * of strain_profiling (profile.rs:3291-3323) as the unit of parallelism.  Contexts are destroyed one by one. */
int ptx_create_multi(const int* devices, int n_devices, int64_t expected_records_per_device, ptx_ctx** out);
int ptx_finalize_multi(ptx_ctx* const* ctxs, int n);

/* ---- introspection ------------------------------------------------------------------ */
/* JSON with per-stage device timings (ms), launch counts, bytes, table sizes. */
int ptx_stats_json(ptx_ctx* ctx, char* buf, size_t cap);
/* Device time (ms, CUDA events on the ctx's stream) of the last ptx_ingest_gaf* + ptx_finalize
 * kernels; number of kernel launches since the last ptx_reset. */
int ptx_timing(ptx_ctx* ctx, double* ingest_ms, double* finalize_ms, int64_t* kernel_launches);

#ifdef __cplusplus
}
#endif
#endif /* PANTAX_GPU_H */
