// sm_100a kernels of the PanTax alignment-to-abundance hot path.
//
//   k_ingest<short|long>   one CTA per tile of GAF text: TMA bulk copy into shared memory, line index, sort by
//                          line length, then one thread per record (short) or one thread per record for the scalar
//                          columns + the whole warp for the walk column (long): column split, integers, walk decode
//                          (rcls.rs:119-146, 237-258), species label, species counts (profile.rs:208-297) ->
//                          record table + CSR walks
//   k_apply<MODE>          one thread per record-table entry: read-id set insert (profile.rs:361-437), node
//                          coverage / trio sums from the CSR walk (profile.rs:787-919); also the replay pass
//   k_count_records        records per 4 KB (+ scan) -> GAF row numbering; first chunk of a ctx / exact redo only
//   k_trio_*               unique-trio table build (profile.rs:658-740)
//   k_cov                  covered bases per node (profile.rs:1018-1023)
//   k_path_len_sum / k_hap_nz / k_depth   per-path / per-hap statistics (profile.rs:980-1016, 1112-1135, 2705-2729)
//   k_ds_*                 cross-rank read-id groups (multi-GPU);  k_flt_*  gaf_filter.rs:22-97
// All HBM-bound integer/byte work: no tensor cores.  See DESIGN.md for the data layout and the
// algorithmic bytes each kernel is measured against.
#include <atomic>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#include "ptx_internal.h"
#include "ptx_fast.cuh"

namespace ptx {

static std::atomic<int64_t> g_launches{0};
int64_t kernel_launch_count() { return g_launches.load(); }
#define PTX_LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

// =====================================================================================
// PTX helpers: mbarrier + TMA 1-D bulk copy (global -> shared), 128-bit CAS
// =====================================================================================
__device__ __forceinline__ uint32_t __smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
// TMA prefetch of a byte range into the L2 (no destination, no completion): the tile a CTA scheduled one wave later will stage
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ ulonglong2 atomic_cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n"
        ".reg .b128 c, v, o;\n"
        "mov.b128 c, {%2, %3};\n"
        "mov.b128 v, {%4, %5};\n"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n"
        "mov.b128 {%0, %1}, o;\n"
        "}\n"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}
__device__ __forceinline__ ulonglong2 ld128(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
// ---- the same accesses with an L2 eviction policy (IngestArgs::pol_keep / pol_stream); atom.cas takes none
__device__ __forceinline__ ulonglong2 ld128_hint(const ulonglong2* p, uint64_t pol) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_v4_hint(const uint4* p, uint64_t pol) {
    uint4 w;
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p), "l"(pol) : "memory");
    return w;
}
__device__ __forceinline__ void red_add64_hint(unsigned long long* p, unsigned long long v, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_or32_hint(uint32_t* p, uint32_t v, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.or.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// 0x80 in every byte of w that equals '\n'
__device__ __forceinline__ uint32_t nl_mask4(uint32_t w) {
    uint32_t y = w ^ 0x0a0a0a0au;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
}
// Does a GAF record start at b[q]?  (not an empty line, not an '@' comment; rcls.rs:123)
__device__ __forceinline__ bool valid_first(const uint8_t* b, uint32_t q) {
    uint8_t c = b[q];
    if (c == '\n' || c == '@') return false;
    if (c == '\r' && b[q + 1] == '\n') return false;
    return true;
}

// =====================================================================================
// K1: records per 4 KB micro-tile (one warp each); tiles of the ingest kernel are runs of micro-tiles
// =====================================================================================
__global__ void __launch_bounds__(256) k_count_records(const uint8_t* __restrict__ text, uint64_t n_bytes, uint32_t n_micro,
                                                       uint32_t* __restrict__ micro_count, unsigned long long* __restrict__ total_slots) {
    const uint32_t mt = blockIdx.x * 8u + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t cnt = 0, slots = 0;
    if (mt < n_micro) {
        const uint8_t* tb = text + (uint64_t)mt * MICRO;
        const uint64_t base = (uint64_t)mt * MICRO;
        constexpr int NJ = (int)(MICRO / 512);
        uint4 q[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = __ldg(reinterpret_cast<const uint4*>(tb + (uint32_t)j * 512u + lane * 16u));  // all loads in flight
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            uint32_t mm = (nl_mask4(q[j].x) >> 7) | (nl_mask4(q[j].y) >> 6) | (nl_mask4(q[j].z) >> 5) | (nl_mask4(q[j].w) >> 4);
            const uint32_t off = (uint32_t)j * 512u + lane * 16u + 1u;
            while (mm) {  // rare: one newline per ~7 pieces
                const uint32_t bit = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t qpos = off + ((bit & 7u) << 2) + (bit >> 3);
                if (base + qpos < n_bytes) {  // a line starts behind this newline (not in the padding)
                    ++slots;
                    cnt += valid_first(tb, qpos) ? 1u : 0u;
                }
            }
        }
        if (mt == 0 && lane == 0 && n_bytes > 0) { ++slots; cnt += valid_first(tb, 0) ? 1u : 0u; }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    slots = __reduce_add_sync(0xffffffffu, slots);
    if (lane == 0 && mt < n_micro) micro_count[mt] = cnt;
    __shared__ uint32_t ws[8];
    if (lane == 0) ws[threadIdx.x >> 5] = slots;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        if (t) atomicAdd(total_slots, (unsigned long long)t);
    }
}


// =====================================================================================
// read-id set (open addressing, 16-byte slots {hash.lo, hash.hi<<32 | epoch<<24 | state}, 128-bit CAS)
//
// Slots carry the EPOCH of the pass that wrote them (1..255, ptx_ctx::ds_epoch): a slot of another epoch is empty.  Starting a new
// pass (ptx_rewind / ptx_reset) is an increment instead of a memset of the whole table (512 MB for 10 M reads at load 0.3, every
// step); the table is cleared for real when the epoch wraps.  Inside a slot the state takes 24 bits: a species index (< 2^24 - 3),
// DS_NONE or DS_MIXED; on the wire (exchange boxes, MIXED lists) entries keep the plain 32-bit state of ptx_core.cuh.
// =====================================================================================
__device__ __forceinline__ uint64_t ds_home(const IdHash& h, uint32_t shift) {
    return ((h.lo ^ ((uint64_t)h.hi << 17)) * 0x9E3779B97F4A7C15ull) >> shift;
}
__device__ __forceinline__ uint32_t ds_enc(uint32_t state, uint32_t ep) { return (ep << 24) | (state & 0xFFFFFFu); }
__device__ __forceinline__ uint32_t ds_dec(uint32_t enc) {
    const uint32_t v = enc & 0xFFFFFFu;
    return v >= 0xFFFFFDu ? (0xFF000000u | v) : v;  // DS_MIXED = 0xFFFFFFFD, DS_NONE = 0xFFFFFFFE
}
__device__ __forceinline__ bool ds_live(const ulonglong2& cur, uint32_t ep) { return (((uint32_t)cur.y) >> 24) == ep; }
__device__ __forceinline__ bool ds_same(const ulonglong2& a, const ulonglong2& b) { return a.x == b.x && a.y == b.y; }

// Insert-or-merge of one id with state `st` (a species, or DS_NONE for a row that is not coverage-eligible; DS_MIXED from the
// wire): profile.rs:369-378 (uniqueness over all non-U rows) + :406-437 (species set per id group over eligible rows only).
// Returns through flags[0] "id seen before", flags[1] "a group became mixed".
__device__ __forceinline__ void ds_upsert(ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, uint64_t lo, uint32_t hi, uint32_t st,
                                          uint32_t* flags, uint64_t pol, bool count_repeat) {
    const uint64_t hi_part = (uint64_t)hi << 32;
    const ulonglong2 mine = make_ulonglong2(lo, hi_part | ds_enc(st, ep));
    IdHash h;
    h.lo = lo;
    h.hi = hi;
    uint64_t i = ds_home(h, shift);
    for (;;) {
        ulonglong2 cur = ld128_hint(slots + i, pol);
        if (!ds_live(cur, ep)) {  // empty (never written, or left by an earlier pass): claim it
            const ulonglong2 prev = atomic_cas128(slots + i, cur, mine);
            if (ds_same(prev, cur)) return;  // inserted
            cur = prev;                       // somebody else claimed it in this pass
            if (!ds_live(cur, ep)) continue;  // (cannot happen: a slot only changes into the current epoch; look again)
        }
        if (cur.x == lo && (cur.y >> 32) == (uint64_t)hi) {  // same read id seen before in this pass
            if (count_repeat && flags[0] == 0u) atomicOr(flags + 0, 1u);
            for (;;) {
                const uint32_t old = ds_dec((uint32_t)cur.y);
                uint32_t nst;
                if (old == DS_NONE) nst = st;
                else if (st == DS_NONE || old == st || old == DS_MIXED) return;
                else nst = DS_MIXED;
                if (nst == old) return;
                const ulonglong2 want = make_ulonglong2(cur.x, hi_part | ds_enc(nst, ep));
                const ulonglong2 prev = atomic_cas128(slots + i, cur, want);
                if (ds_same(prev, cur)) {
                    if (nst == DS_MIXED) atomicOr(flags + 1, 1u);
                    return;
                }
                cur = prev;
            }
        }
        i = (i + 1) & mask;
    }
}
__device__ __forceinline__ void ds_insert(ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, const IdHash& h, bool eligible,
                                          uint32_t label, uint32_t* flags, uint64_t pol) {
    ds_upsert(slots, shift, mask, ep, h.lo, h.hi, eligible ? label : DS_NONE, flags, pol, true);
}

__device__ __forceinline__ uint32_t ds_lookup(const ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, const IdHash& h, uint64_t pol) {
    uint64_t i = ds_home(h, shift);
    for (;;) {
        ulonglong2 cur = ld128_hint(slots + i, pol);
        if (!ds_live(cur, ep)) return DS_NONE;
        if (cur.x == h.lo && (cur.y >> 32) == (uint64_t)h.hi) return ds_dec((uint32_t)cur.y);
        i = (i + 1) & mask;
    }
}

// the live slots of the old table move to a new (zeroed) one, keeping their epoch
__global__ void __launch_bounds__(256) k_ds_rehash(const ulonglong2* __restrict__ old_slots, uint64_t old_cap, ulonglong2* new_slots,
                                                   uint32_t new_shift, uint64_t new_mask, uint32_t ep) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < old_cap; k += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 e = old_slots[k];
        if (!ds_live(e, ep)) continue;
        IdHash h;
        h.lo = e.x;
        h.hi = (uint32_t)(e.y >> 32);
        uint64_t i = ds_home(h, new_shift);
        for (;;) {
            const ulonglong2 cur = ld128(new_slots + i);
            if (!ds_live(cur, ep)) {
                const ulonglong2 prev = atomic_cas128(new_slots + i, cur, e);
                if (ds_same(prev, cur)) break;
            }
            i = (i + 1) & new_mask;
        }
    }
}

// ---- cross-rank id groups (multi-GPU): an id may repeat on ANOTHER rank's read batch ----------------
// Every id-set entry is routed to the rank that owns its hash (owner = f(hash) % P); the owner merges the
// per-rank states, and the hashes whose merged state is DS_MIXED are broadcast back.
__device__ __forceinline__ uint32_t ds_owner(const ulonglong2& e, uint32_t P) {
    return ((uint32_t)(e.y >> 32) ^ (uint32_t)(e.x >> 40)) % P;
}
// Owner side: merge the entries received from peer blockIdx.y (inbox + off[peer], cnt[peer] entries; wire format) into this rank's
// id set - the same insert-or-merge as ds_insert, with the sender's state instead of a single record's.
__global__ void __launch_bounds__(256) k_ds_merge_boxes(const ulonglong2* __restrict__ inbox, const unsigned long long* __restrict__ off,
                                                        const unsigned long long* __restrict__ cnt, ulonglong2* slots, uint32_t shift,
                                                        uint64_t mask, uint32_t ep, uint32_t* flags) {
    const ulonglong2* box = inbox + off[blockIdx.y];
    const uint64_t n = cnt[blockIdx.y];
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 e = box[k];
        ds_upsert(slots, shift, mask, ep, e.x, (uint32_t)(e.y >> 32), (uint32_t)e.y, flags, 0x1000000000000000ull, true);
    }
}
// the MIXED ids of this rank's set, in wire format
__global__ void __launch_bounds__(256) k_ds_collect_mixed(const ulonglong2* __restrict__ slots, uint64_t cap, uint32_t ep, unsigned long long* cursor,
                                                          ulonglong2* __restrict__ out, uint64_t out_cap) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cap; k += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 e = slots[k];
        if (!ds_live(e, ep) || ds_dec((uint32_t)e.y) != DS_MIXED) continue;
        const unsigned long long j = atomicAdd(cursor, 1ull);
        if (out && j < out_cap) out[j] = make_ulonglong2(e.x, (e.y & 0xFFFFFFFF00000000ull) | DS_MIXED);
    }
}
// MIXED ids of all ranks (wire format; zero entries are padding of the all-gather): mark them here - an id this rank does not own is
// inserted as MIXED so that the keep-mask lookup of this rank's reads with that id finds it
__global__ void __launch_bounds__(256) k_ds_apply_mixed(const ulonglong2* __restrict__ in, uint64_t n, ulonglong2* slots, uint32_t shift,
                                                        uint64_t mask, uint32_t ep, uint32_t* scratch_flags) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 e = in[k];
        if (e.x == 0ull && e.y == 0ull) continue;
        ds_upsert(slots, shift, mask, ep, e.x, (uint32_t)(e.y >> 32), DS_MIXED, scratch_flags, 0x1000000000000000ull, false);
    }
}

// =====================================================================================
// k_ingest: GAF text -> species counts, record table, CSR walks   (k_apply, further down, consumes them)
// =====================================================================================
struct DevSink {
    const IngestArgs& a;
    // one 16-byte gather: {len, flags, bit_off}.  The flags word is updated by atomics while we read it with a
    // plain load: a stale "not full" only costs a redundant atomicOr.
    __device__ __forceinline__ NodeInfo info(uint32_t g) const {
        const uint4 w = ld_v4_hint(a.ninfo + g, a.pol_keep);
        NodeInfo ni;
        ni.len = w.x;
        ni.flags = w.y;
        ni.bit_off = ((uint64_t)w.w << 32) | w.z;
        return ni;
    }
    __device__ __forceinline__ void add_bases(uint32_t g, int64_t v) { red_add64_hint(a.bases + g, (unsigned long long)v, a.pol_keep); }
    __device__ __forceinline__ void set_bits(uint32_t g, const NodeInfo& ni, int64_t lo, int64_t hi) {
        if (lo == 0 && hi == (int64_t)ni.len) {  // whole node: one flag, set once
            if (!(ni.flags & NI_FULL)) red_or32_hint(&a.ninfo[g].y, NI_FULL, a.pol_keep);
            return;
        }
        if (ni.flags & NI_FULL) return;  // already fully covered: partial intervals add nothing
        const uint64_t b0 = ni.bit_off + (uint64_t)lo, b1 = ni.bit_off + (uint64_t)hi - 1;  // inclusive last bit
        const uint64_t w0 = b0 >> 5, w1 = b1 >> 5;
        const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u), m1 = 0xFFFFFFFFu >> (31u - (uint32_t)(b1 & 31u));
        if (w0 == w1) {
            red_or32_hint(a.bits + w0, m0 & m1, a.pol_keep);
        } else {
            red_or32_hint(a.bits + w0, m0, a.pol_keep);
            for (uint64_t w = w0 + 1; w < w1; ++w) red_or32_hint(a.bits + w, 0xFFFFFFFFu, a.pol_keep);
            red_or32_hint(a.bits + w1, m1, a.pol_keep);
        }
    }
    __device__ __forceinline__ void trio(uint32_t x, uint32_t y, uint32_t z, int64_t s, uint32_t y_flags) {
        if (a.tt == nullptr) return;
        const uint32_t lo = x < z ? x : z, hi = x < z ? z : x;  // profile.rs:672-678 / :902-904
        if (!(y_flags & trio_sig_bit(lo, hi))) return;           // no unique trio around y has these neighbours
        uint32_t i = trio_hash(lo, y, hi) & a.tt_mask;
        for (;;) {
            uint4 e = __ldg(a.tt + i);
            if (e.x == TT_EMPTY) return;
            if (e.x == lo && e.y == y && e.z == hi) {
                atomicAdd(a.trio_bases + e.w, (unsigned long long)s);
                return;
            }
            i = (i + 1) & a.tt_mask;
        }
    }
    __device__ __forceinline__ void error_start_gt_len(uint32_t label) { atomicOr(a.err + label, 1u); }
};

// ---- the other scatter variants of north_star stage 2: same sink, add_bases replaced (profiles/r2_scatter_bakeoff.md)
// 1: the lanes of the warp that add to the same node in this step add once (match.any + shuffles among the group)
struct WarpAggSink : DevSink {
    __device__ __forceinline__ void add_bases(uint32_t g, int64_t v) {
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, g);
        const uint32_t lane = threadIdx.x & 31u;
        if ((peers & (peers - 1u)) == 0u) { atomicAdd(a.bases + g, (unsigned long long)v); return; }
        const int leader = __ffs(peers) - 1;
        long long sum = 0;
        for (unsigned m = peers; m; m &= m - 1u) sum += __shfl_sync(peers, (long long)v, __ffs(m) - 1);  // same trip count for the whole group
        if ((int)lane == leader) atomicAdd(a.bases + g, (unsigned long long)sum);
    }
};
// 2: a per-CTA shared-memory table keyed by node absorbs the nodes a CTA meets more than once ("hot" node ranges: abundant
// species, short graphs); a slot taken by another node falls back to the global RED.  Flushed once per CTA.
constexpr uint32_t SC_SLOTS_LOG2 = 11, SC_SLOTS = 1u << SC_SLOTS_LOG2;
struct SmemSink : DevSink {
    uint32_t* tag;
    unsigned long long* sum;
    __device__ __forceinline__ void add_bases(uint32_t g, int64_t v) {
        const uint32_t s = (g * 0x9E3779B1u) >> (32u - SC_SLOTS_LOG2);
        const uint32_t old = atomicCAS(tag + s, 0xFFFFFFFFu, g);
        if (old == 0xFFFFFFFFu || old == g) atomicAdd(sum + s, (unsigned long long)v);
        else atomicAdd(a.bases + g, (unsigned long long)v);
    }
};
// 3: nothing is added here - the (node, bases) pair goes to the record's CSR slots; launch_scatter_sorted sorts and reduces them
struct PairSink : DevSink {
    uint32_t pos;
    __device__ __forceinline__ void add_bases(uint32_t g, int64_t v) {
        a.pair_key[pos] = g;
        a.pair_val[pos] = (unsigned long long)v;
        ++pos;
    }
};
__global__ void __launch_bounds__(256) k_add_runs(const uint32_t* __restrict__ key, const unsigned long long* __restrict__ sum,
                                                  const unsigned long long* __restrict__ n_runs, unsigned long long* bases) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n_runs && key[i] != 0xFFFFFFFFu) bases[key[i]] += sum[i];  // keys are unique: no atomics
}

// ---- warp-cooperative decode of the walk column (long reads: a tile holds few lines, each with a walk of tens
// to hundreds of nodes).  The 32 lanes take 4 bytes each of a 128-byte window of the column; byte classes come
// from SWAR masks, node starts from ballots, every lane that owns a start converts that digit run.
__device__ __forceinline__ uint32_t digit_mask4(uint32_t w) {  // 0x80 in every byte that is '0'..'9'
    const uint32_t t = w ^ 0x30303030u;
    const uint32_t bad = (t & 0xF0F0F0F0u) | (((t & 0x0F0F0F0Fu) + 0x06060606u) & 0x10101010u);
    return ~(((bad & 0x7f7f7f7fu) + 0x7f7f7f7fu) | bad) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nib4(uint32_t m) { return ((m >> 7) * 0x10204080u) >> 28; }  // bit i = byte i

struct CoopCount { uint32_t W, end; int term; bool beyond; };
// Column 6 starting at stage[p6]: number of digit runs, position and class of its terminator ('\t', '\n' or the
// '\r' of "\r\n", as term_at), `beyond` if it is not inside the window (lim = staged bytes, sentinel '\n' behind).
__device__ __forceinline__ CoopCount coop_walk_count(const uint8_t* stage, uint32_t p6, uint32_t lim, uint32_t lane) {
    uint32_t pos = p6 & ~3u, W = 0, carry = 0;
    for (;;) {
        const uint32_t my = pos + 4u * lane;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + my);
        uint32_t dn = nib4(digit_mask4(w));
        uint32_t tn = nib4(eq_mask4(w, 0x09090909u) | eq_mask4(w, 0x0a0a0a0au));
        if (my < p6) { dn &= 0xFu << (p6 - my); tn &= 0xFu << (p6 - my); }  // lane 0 of the first window only
        const unsigned tb = __ballot_sync(0xffffffffu, tn != 0u);
        uint32_t endpos = 0;
        if (tb) {
            const uint32_t L = __ffs(tb) - 1;
            const uint32_t bit = __ffs(__shfl_sync(0xffffffffu, tn, L)) - 1;
            endpos = pos + 4u * L + bit;
            if (lane > L) dn = 0; else if (lane == L) dn &= (1u << bit) - 1u;
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, dn >> 3, 1);
        const uint32_t sn = dn & ~((dn << 1) | (lane == 0 ? carry : (prev & 1u))) & 0xFu;  // first digit of a run
        const uint32_t c = __popc(sn);
        W += __popc(__ballot_sync(0xffffffffu, c >= 1u)) + __popc(__ballot_sync(0xffffffffu, c == 2u));
        if (tb) {
            CoopCount r;
            r.W = W;
            r.beyond = endpos + 1u >= lim;
            r.term = stage[endpos] == '\t' ? T_TAB : T_EOL;
            if (r.term == T_EOL && endpos > p6 && stage[endpos - 1] == '\r') --endpos;
            r.end = endpos;
            return r;
        }
        carry = __shfl_sync(0xffffffffu, dn >> 3, 31) & 1u;
        pos += 128u;
    }
}

struct CoopWalk { uint32_t W, end, vmin, vmax; int term; bool beyond, monotone, big; };
// One sweep over column 6 starting at stage[p6]: finds its terminator (as coop_walk_count) and decodes the digit
// runs before it to dst[0..W) in walk order (runs of <= 9 digits; `big` reports a longer one, whose record is then
// redone by the scalar 64-bit scanner), with min, max and strict monotonicity.  dst must have room for every run
// the line can hold ((bytes from p6 to the end of the line) / 2 + 1).
__device__ __forceinline__ CoopWalk coop_walk_sweep(const uint8_t* stage, uint32_t p6, uint32_t lim, uint32_t* dst, uint32_t lane) {
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t Wb = 0, carry = 0, carry_last = 0, mn = 0xFFFFFFFFu, mx = 0, endpos = 0;
    bool inc = true, dec = true, big = false;
    for (uint32_t pos = p6 & ~3u;; pos += 128u) {
        const uint32_t my = pos + 4u * lane;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(stage + my);
        uint32_t dn = nib4(digit_mask4(w));
        uint32_t tn = nib4(eq_mask4(w, 0x09090909u) | eq_mask4(w, 0x0a0a0a0au));
        if (my < p6) { dn &= 0xFu << (p6 - my); tn &= 0xFu << (p6 - my); }  // lane 0 of the first window only
        const unsigned tb = __ballot_sync(0xffffffffu, tn != 0u);
        if (tb) {  // the column ends in this window: nothing behind the terminator counts
            const uint32_t L = __ffs(tb) - 1;
            const uint32_t bit = __ffs(__shfl_sync(0xffffffffu, tn, L)) - 1;
            endpos = pos + 4u * L + bit;
            if (lane > L) dn = 0; else if (lane == L) dn &= (1u << bit) - 1u;
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, dn >> 3, 1);
        uint32_t sn = dn & ~((dn << 1) | (lane == 0 ? carry : (prev & 1u))) & 0xFu;
        const uint32_t c = __popc(sn);  // 0, 1 or 2 node ids start in these 4 bytes
        const unsigned b1 = __ballot_sync(0xffffffffu, c >= 1u), b2 = __ballot_sync(0xffffffffu, c == 2u);
        const uint32_t ord = Wb + __popc(b1 & lt) + __popc(b2 & lt);
        uint32_t first_v = 0, last_v = 0;
        for (uint32_t k = 0; k < c; ++k) {
            uint32_t qb = my + __ffs(sn) - 1u;
            sn &= sn - 1u;
            uint32_t d = (uint32_t)stage[qb] - (uint32_t)'0', v = 0, nd = 0;
            while (d <= 9u && nd < 9u) {
                v = v * 10u + d;
                ++nd;
                d = (uint32_t)stage[++qb] - (uint32_t)'0';
            }
            if (d <= 9u) big = true;
            dst[ord + k] = v;
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
            if (k == 0) first_v = v;
            else { if (v <= last_v) inc = false; if (v >= last_v) dec = false; }
            last_v = v;
        }
        // the id before this lane's first one: last id of the nearest lower lane that has any, else of the previous window
        const unsigned below = b1 & lt;
        const uint32_t pv_lane = __shfl_sync(0xffffffffu, last_v, below ? 31u - (uint32_t)__clz(below) : 0u);
        if (c && (below || Wb)) {
            const uint32_t pv = below ? pv_lane : carry_last;
            if (first_v <= pv) inc = false;
            if (first_v >= pv) dec = false;
        }
        const uint32_t top = __shfl_sync(0xffffffffu, last_v, b1 ? 31u - (uint32_t)__clz(b1) : 0u);
        if (b1) carry_last = top;
        Wb += __popc(b1) + __popc(b2);
        if (tb) break;
        carry = __shfl_sync(0xffffffffu, dn >> 3, 31) & 1u;
    }
    CoopWalk r;
    r.W = Wb;
    r.beyond = endpos + 1u >= lim;
    r.term = stage[endpos] == '\t' ? T_TAB : T_EOL;
    if (r.term == T_EOL && endpos > p6 && stage[endpos - 1] == '\r') --endpos;
    r.end = endpos;
    r.vmin = __reduce_min_sync(0xffffffffu, mn);
    r.vmax = __reduce_max_sync(0xffffffffu, mx);
    r.monotone = __all_sync(0xffffffffu, inc) || __all_sync(0xffffffffu, dec);
    r.big = __any_sync(0xffffffffu, big);
    return r;
}

// rare path: the record did not fit in the staged window - parse it from global memory.
// Works on its own RecParse so that the caller's stays in registers.
__device__ __noinline__ void parse_record_global(const uint8_t* b, uint32_t p, uint32_t lim, RecParse* out) {
    RecParse r;
    parse_record(b, p, lim, r, 1u << (threadIdx.x & 31u), nullptr, 0, 0);
    *out = r;
}

template <bool LONG>
__global__ void __launch_bounds__(INGEST_THREADS, LONG ? 3 : 4) k_ingest(const IngestArgs a) {
    constexpr int MODE = MODE_CLASSIFY;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile_bytes = a.tile_bytes;          // multiple of 4096 for this kernel
    const uint32_t rows = tile_bytes / ((INGEST_THREADS / 32u) * 512u);  // 1..8 rows of 512 B per warp
    const uint32_t stage_bytes = tile_bytes + OVER;
    uint8_t* stage = smem;  // tile_bytes + OVER, then 16 sentinel '\n' (the column scanners stop at a newline)
    uint32_t* stash = reinterpret_cast<uint32_t*>(smem + stage_bytes + 16);                  // [STASH_CAP][INGEST_THREADS]
    uint16_t* rec_start = reinterpret_cast<uint16_t*>(stash + STASH_CAP * INGEST_THREADS);  // [REC_CAP] record starts
    uint16_t* rec_tmp = rec_start + REC_CAP;                                                 // [REC_CAP] (bin, rank in bin)
    uint16_t* order = rec_tmp + REC_CAP;                                                     // [REC_CAP] records by line length
    uint16_t* inv_pre = order + REC_CAP;                                                     // [REC_CAP] invalid line slots before slot k
    // species-count accumulators of this tile, only when S > 1: a small open-addressed table keyed by label.
    // Abundant species (the contended global counters) are absorbed here and flushed once per tile.
    uint32_t* hkey = reinterpret_cast<uint32_t*>(inv_pre + REC_CAP);                         // [HIST_SLOTS]
    uint32_t* hval = hkey + HIST_SLOTS;                                                      // [HIST_SLOTS][4]: n, less_multi, uniq, sum qlen
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t warp_tot[INGEST_THREADS / 32];
    __shared__ uint32_t bin_cnt[64];
    __shared__ uint32_t inv_flag, inv_tot_s, slot_base_s;

    const uint64_t t0 = (uint64_t)blockIdx.x * tile_bytes;
    const uint8_t* gtile = a.text + t0;

    // ---- stage the tile: one TMA bulk copy, completion on an mbarrier
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&mbar, stage_bytes);
        bulk_g2s(stage, gtile, stage_bytes, &mbar);
    }
    if (tid < 4) reinterpret_cast<uint32_t*>(stage + stage_bytes)[tid] = 0x0a0a0a0au;  // sentinel behind the window
    const bool hist_smem = a.ranges.S > 1;
    if (hist_smem)
        for (uint32_t i = tid; i < HIST_SLOTS * 5u; i += INGEST_THREADS) hkey[i] = i < HIST_SLOTS ? LABEL_U : 0u;
    const uint32_t* sstart = a.ranges.sstart;  // 4 B x S, contiguous: stays in L1 across the binary searches
    mbar_wait(&mbar, 0);

    const uint32_t wbase_byte = warp * rows * 512u;
    const bool last_tile = t0 + tile_bytes + 1u >= a.n_bytes;
    const bool single_pass = a.micro_base == nullptr;  // no count pass: rows are numbered per tile (tile_info + row_key)
    const uint32_t rec_base = single_pass ? 0u : (uint32_t)a.micro_base[(uint64_t)blockIdx.x * rows];
    const uint32_t glim = (uint32_t)min((uint64_t)0xFFFF0000ull, a.padded_bytes - t0);
    const RangesView& R = a.ranges;
    __syncthreads();  // sentinel visible

    uint32_t n_rec = 0;       // line slots of the tile: every line start behind an owned newline (valid record or not)
    uint32_t valid_prev = 0;  // valid records in the previous rounds of this tile
    for (uint32_t round = 0; round == 0 || round < n_rec; round += REC_CAP) {
        // ---- line starts: warp w scans stage[w*rows*512, +rows*512) as `rows` rows of 32 x 16 B.  Per 16-byte
        // piece the 16 newline flags are packed into one word (bit = word + 8*byte); no per-newline loop.
        // (Recomputed in the rare extra rounds of a tile with more than REC_CAP lines, so that the per-row
        // words do not stay live in registers while the records are processed.)
        {
            uint32_t mmv[MAX_ROWS];
#pragma unroll
            for (int j = 0; j < (int)MAX_ROWS; ++j) {
                uint32_t mm = 0;
                if ((uint32_t)j < rows) {
                    const uint4 q = *reinterpret_cast<const uint4*>(stage + wbase_byte + (uint32_t)j * 512u + lane * 16u);
                    mm = (nl_mask4(q.x) >> 7) | (nl_mask4(q.y) >> 6) | (nl_mask4(q.z) >> 5) | (nl_mask4(q.w) >> 4);
                    if (mm && last_tile) {  // a line start at or beyond the end of the text is padding, not a line
                        const uint64_t off = t0 + wbase_byte + (uint32_t)j * 512u + lane * 16u + 1u;
                        for (uint32_t pos = 0; pos < 16; ++pos)
                            if (off + pos >= a.n_bytes) mm &= ~(1u << ((pos >> 2) + 8u * (pos & 3u)));
                    }
                }
                mmv[j] = mm;
            }
            const uint32_t first_slot = (blockIdx.x == 0 && tid == 0) ? 1u : 0u;  // the line at text[0] (behind the PRE padding)
            // exclusive position of (row j, lane) inside the warp, rows first
            uint32_t pre[MAX_ROWS];
            uint32_t run = 0;
#pragma unroll
            for (int j = 0; j < (int)MAX_ROWS; ++j) {
                const uint32_t cj = __popc(mmv[j]) + (j == 0 ? first_slot : 0u);
                uint32_t x = cj;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                pre[j] = run + x - cj;
                run += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) warp_tot[warp] = run;
            if (tid == 0) inv_flag = 0;
            __syncthreads();
            uint32_t my_base = 0;
            n_rec = 0;
#pragma unroll
            for (int w = 0; w < INGEST_THREADS / 32; ++w) {
                uint32_t t = warp_tot[w];
                if ((uint32_t)w < warp) my_base += t;
                n_rec += t;
            }
            // ---- compact this round's line starts into shared memory
#pragma unroll
            for (int j = 0; j < (int)MAX_ROWS; ++j) {
                const uint32_t mm = mmv[j];
                uint32_t idx = my_base + pre[j];
                if (j == 0 && first_slot) {
                    if (idx >= round && idx < round + REC_CAP) rec_start[idx - round] = 0;
                    ++idx;
                }
                if (mm == 0) continue;
                const uint32_t off = wbase_byte + (uint32_t)j * 512u + lane * 16u + 1u;
                if ((mm & (mm - 1)) == 0) {  // one newline in the piece (lines are longer than 16 bytes)
                    const uint32_t bit = __ffs(mm) - 1;
                    if (idx >= round && idx < round + REC_CAP) rec_start[idx - round] = (uint16_t)(off + ((bit & 7u) << 2) + (bit >> 3));
                } else {  // several very short lines: emit in byte order
                    for (uint32_t pos = 0; pos < 16; ++pos) {
                        if ((mm >> ((pos >> 2) + 8u * (pos & 3u))) & 1u) {
                            if (idx >= round && idx < round + REC_CAP) rec_start[idx - round] = (uint16_t)(off + pos);
                            ++idx;
                        }
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t n_round = min(REC_CAP, n_rec - round);

        // ---- empty lines and '@' comments are line slots but not records: they are rare and only shift the record
        // numbering (labels[]); inv_pre[k] = invalid slots before slot k is built only if the tile has any
        uint32_t inv_total = 0;
        if (MODE & MODE_CLASSIFY) {
            bool inv = false;
            for (uint32_t k = tid; k < n_round; k += INGEST_THREADS) inv |= !valid_first(stage, rec_start[k]);
            if (inv) inv_flag = 1;
            __syncthreads();
            if (inv_flag) {
                if (warp == 0) {
                    uint32_t base = 0;
                    for (uint32_t k0 = 0; k0 < n_round; k0 += 32) {
                        const uint32_t k = k0 + lane;
                        const bool bad = k < n_round && !valid_first(stage, rec_start[k]);
                        const unsigned bm = __ballot_sync(0xffffffffu, bad);
                        if (k < n_round) inv_pre[k] = (uint16_t)(base + __popc(bm & ((1u << lane) - 1u)));
                        base += __popc(bm);
                    }
                    if (lane == 0) inv_tot_s = base;
                }
                __syncthreads();
                inv_total = inv_tot_s;
            }
        }

        // ---- order the round's records by line length (a proxy for the walk length: 4-byte bins) so that the
        // lanes of a warp carry walks of similar length; the lock-step node loops then idle much less
        if (tid < 64) bin_cnt[tid] = 0;
        if (tid == 64) {
            uint32_t sb = atomicAdd(a.cursors + 0, n_round);  // this round's entries in the record table
            if (single_pass) {
                // the table was sized from an estimate, and row numbering per tile needs all lines of a tile in one round
                if (sb + n_round > a.slots_cap || n_rec > REC_CAP) {
                    atomicOr(a.cursors + 3, 1u);  // the host redoes the chunk with the count pass
                    sb = 0xFFFFFFFFu;
                } else {
                    a.tile_info[blockIdx.x] = make_uint4(sb, n_round, n_round - inv_total, 0u);
                    atomicAdd(a.cursors + 2, n_round - inv_total);
                }
            }
            slot_base_s = sb;
        }
        __syncthreads();
        if (slot_base_s == 0xFFFFFFFFu) return;  // uniform: read after the barrier
        for (uint32_t k = tid; k < n_round; k += INGEST_THREADS) {
            const uint32_t s0 = rec_start[k];
            const uint32_t e0 = (k + 1 < n_round) ? rec_start[k + 1] : s0 + 112u;
            const uint32_t len = e0 - s0;
            const uint32_t bin = len < 64u ? 0u : min((len - 64u) >> 2, 63u);
            rec_tmp[k] = (uint16_t)((bin << 10) | atomicAdd(&bin_cnt[bin], 1u));
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 64 bin counts
            const uint32_t c0 = bin_cnt[2 * lane], c1 = bin_cnt[2 * lane + 1];
            uint32_t x = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            bin_cnt[2 * lane] = x - c0 - c1;
            bin_cnt[2 * lane + 1] = x - c1;
        }
        __syncthreads();
        for (uint32_t k = tid; k < n_round; k += INGEST_THREADS) {
            const uint32_t t = rec_tmp[k];
            order[bin_cnt[t >> 10] + (t & 1023u)] = (uint16_t)k;
        }
        __syncthreads();

        // ---- one thread per record; the lanes of a warp move through the columns in lock-step
        for (uint32_t k0 = 0; k0 < n_round; k0 += INGEST_THREADS) {
            // LONG: records are dealt round-robin to the warps (a tile holds far fewer lines than threads)
            const uint32_t q = LONG ? k0 + lane * (INGEST_THREADS / 32u) + warp : k0 + tid;
            const bool slot = q < n_round;
            const uint32_t k = slot ? order[q] : 0u;
            const uint32_t p = slot ? rec_start[k] : 0u;
            const bool has = slot && valid_first(stage, p);  // not an empty line / '@' comment (rcls.rs:123)
            const uint32_t pmask = __ballot_sync(0xffffffffu, has);
            RecParse r;
            r.W = 0; r.mapq = NULL_I64; r.qlen = NULL_I64; r.stashed = false;
            uint32_t label = LABEL_U;
            const uint8_t* b = stage;
            uint32_t node_off = 0;
            uint32_t w_res = 0;  // LONG: CSR slots this record already holds at node_off (reserved from its line length)
            bool slow = false;  // LONG: this record goes through the scalar parser on the global copy of its line
            if constexpr (LONG) {
                // columns 1-5 per thread, column 6 by the whole warp one record at a time, columns 7-12 per thread
                r.qlen = r.c7 = r.c8 = r.c9 = r.mapq = NULL_I64;
                r.vmin = INT64_MAX; r.vmax = -1; r.path_null = true; r.monotone = true;
                uint32_t pp = p;
                int st = T_EOL;
                if (has) {
                    st = parse_head(stage, pp, r, pmask);
                    if (pp + 1u >= stage_bytes) slow = true;
                }
                __syncwarp();
                r.path_pos = r.path_end = pp;
                bool want6 = has && !slow && st == T_TAB;
                // CSR slots: every digit run takes at least two bytes of the line, so (bytes from column 6 to the end
                // of the line) / 2 + 1 slots always suffice - no counting sweep.  The end of the LAST line of the tile
                // is not known (it lies beyond the line index): that one record is counted first.
                uint32_t w_up = 0, end6 = pp;
                int term6 = st;
                const bool last_line = k + 1u >= n_round;
                if (want6 && !last_line) w_up = ((uint32_t)rec_start[k + 1] - pp) / 2u + 1u;
                for (unsigned m = __ballot_sync(0xffffffffu, want6 && last_line); m; m &= m - 1u) {
                    const int rl = __ffs(m) - 1;
                    const CoopCount cc = coop_walk_count(stage, __shfl_sync(0xffffffffu, pp, rl), stage_bytes, lane);
                    if ((int)lane == rl) { w_up = cc.W; end6 = cc.end; term6 = cc.term; slow = cc.beyond; }
                }
                want6 = want6 && !slow;
                {  // node slots of every record decoded here (eligible or not)
                    const uint32_t wa = want6 ? w_up : 0u;
                    w_res = wa;
                    uint32_t x = wa;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                        if (lane >= (uint32_t)d) x += y;
                    }
                    const uint32_t wtot = __shfl_sync(0xffffffffu, x, 31);
                    uint32_t nbase = 0;
                    if (lane == 0 && wtot) nbase = atomicAdd(a.cursors + 1, wtot);
                    nbase = __shfl_sync(0xffffffffu, nbase, 0);
                    node_off = nbase + x - wa;
                }
                for (unsigned m = __ballot_sync(0xffffffffu, want6 && w_up != 0u); m; m &= m - 1u) {
                    const int rl = __ffs(m) - 1;
                    const CoopWalk cw = coop_walk_sweep(stage, __shfl_sync(0xffffffffu, pp, rl), stage_bytes,
                                                        a.nodes + __shfl_sync(0xffffffffu, node_off, rl), lane);
                    if ((int)lane == rl) {
                        r.W = cw.W;
                        end6 = cw.end;
                        term6 = cw.term;
                        if (cw.W) { r.vmin = (int64_t)cw.vmin; r.vmax = (int64_t)cw.vmax; }
                        r.monotone = cw.monotone;
                        if (cw.big || cw.beyond) slow = true;  // 10+ digit run (generic 64-bit scan) / line leaves the window
                    }
                }
                want6 = want6 && !slow;
                const uint32_t tmask = __ballot_sync(0xffffffffu, want6);
                if (want6) {
                    r.path_end = end6;
                    r.path_null = (end6 - pp == 1u) && (stage[pp] == '*');
                    uint32_t pt = end6;
                    if (term6 == T_TAB) ++pt;
                    parse_tail(stage, pt, term6, r, tmask);
                    if (pt + 1u >= stage_bytes) slow = true;
                }
                __syncwarp();
                if (has && slow) {
                    RecParse tmp;
                    b = gtile;
                    parse_record_global(gtile, p, glim, &tmp);
                    r = tmp;
                }
                __syncwarp();
            } else {
                if (has) {
                    if (!parse_record(stage, p, stage_bytes, r, pmask, stash + tid, INGEST_THREADS, STASH_CAP)) {
                        // the columns run past the staged window (at most one record per tile; long lines only)
                        RecParse tmp;
                        b = gtile;
                        parse_record_global(gtile, p, glim, &tmp);
                        r = tmp;
                    }
                }
            }
            if (has) {
                const uint32_t row = rec_base + valid_prev + k - (inv_total ? (uint32_t)inv_pre[k] : 0u);  // GAF row within the chunk
                if (a.labels_in) {  // strain-only resume: the species column of reads_classification.tsv
                    label = a.labels_in[row];
                    if (label != LABEL_U && r.W && (r.vmin < R.start[label] || r.vmax > R.end[label])) {
                        atomicOr(a.flags + 3, 1u);  // the walk leaves the species graph: reported by ptx_finalize
                        label = LABEL_U;
                    }
                } else {
                    label = classify(R, r.W ? r.vmin : -1, r.W ? r.vmax : -1, sstart);
                }
                if (!single_pass) a.labels[row] = label;
            }
            __syncwarp();
            if (MODE & MODE_CLASSIFY) {
                // ---- species counts (profile.rs:219-232, 264-277), warp-aggregated when the warp is one species
                const bool cnt = has && label != LABEL_U;
                const unsigned mm = __ballot_sync(0xffffffffu, cnt);
                if (mm) {
                    const int leader = __ffs(mm) - 1;
                    const uint32_t lab0 = __shfl_sync(0xffffffffu, label, leader);
                    const bool uniform = __all_sync(0xffffffffu, !cnt || label == lab0);
                    const bool mq_ok = cnt && r.mapq != NULL_I64 && r.mapq >= 3 && r.mapq <= 60;
                    const uint32_t lm = mq_ok ? 1u : 0u;
                    const uint32_t uq = (mq_ok && r.mapq == 60) ? 1u : 0u;
                    const unsigned long long ql = (cnt && r.qlen != NULL_I64) ? (unsigned long long)r.qlen : 0ull;
                    if (uniform) {
                        uint32_t n1 = __popc(mm);
                        uint32_t n3 = __reduce_add_sync(0xffffffffu, lm);
                        uint32_t n4 = __reduce_add_sync(0xffffffffu, uq);
                        unsigned long long s = ql;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                        if ((int)lane == leader) {
                            unsigned long long* hp = a.hist + 4ull * lab0;
                            atomicAdd(hp + 0, (unsigned long long)n1);
                            atomicAdd(hp + 1, s);
                            if (n3) atomicAdd(hp + 2, (unsigned long long)n3);
                            if (n4) atomicAdd(hp + 3, (unsigned long long)n4);
                        }
                    } else if (cnt) {
                        bool done = false;
                        if (ql < (1ull << 16)) {  // a tile holds < 64 K lines: the 32-bit sum cannot wrap
                            uint32_t hs = (label * 0x9E3779B1u) >> (32 - HIST_SLOTS_LOG2);
                            for (int t = 0; t < 4 && !done; ++t) {
                                const uint32_t old = atomicCAS(&hkey[hs], LABEL_U, label);
                                if (old == LABEL_U || old == label) {
                                    uint32_t* hv = hval + 4u * hs;
                                    atomicAdd(hv + 0, 1u);
                                    if (lm) atomicAdd(hv + 1, 1u);
                                    if (uq) atomicAdd(hv + 2, 1u);
                                    if (ql) atomicAdd(hv + 3, (uint32_t)ql);
                                    done = true;
                                }
                                hs = (hs + 1u) & (HIST_SLOTS - 1u);
                            }
                        }
                        if (!done) {  // table full around this slot (rare species) or a huge read length
                            unsigned long long* hp = a.hist + 4ull * label;
                            atomicAdd(hp + 0, 1ull);
                            atomicAdd(hp + 1, ql);
                            if (lm) atomicAdd(hp + 2, 1ull);
                            if (uq) atomicAdd(hp + 3, 1ull);
                        }
                    }
                }
            }
            // ---- emit the record for k_apply: id hash, label, alignment interval and the walk as CSR node ids.
            // Table entries follow the length-sorted thread order, so k_apply's warps also see similar walks.
            const bool labelled = has && label != LABEL_U;
            const bool eligible = labelled && !r.path_null && r.c7 != NULL_I64 && r.c8 != NULL_I64 && r.c9 != NULL_I64;  // profile.rs:380-399
            const uint32_t wcnt = eligible ? r.W : 0u;
            if constexpr (LONG) {
                // the cooperative decode already wrote the walk at node_off.  The scalar fallback writes into the slots the record
                // reserved from its line length (every digit run takes two bytes: they always suffice) and takes new ones only if
                // it reserved none (its line left the window before column 6) - never both, so that the slots of a chunk stay
                // within text bytes / 2 + one per line (chunk_nodes_ensure) whatever the ids look like
                if (slow && wcnt > w_res) node_off = atomicAdd(a.cursors + 1, wcnt);
            } else {
                uint32_t x = wcnt;  // node slots: warp scan, one atomicAdd per warp on the chunk's node cursor
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                const uint32_t wtot = __shfl_sync(0xffffffffu, x, 31);
                uint32_t nbase = 0;
                if (lane == 0 && wtot) nbase = atomicAdd(a.cursors + 1, wtot);
                nbase = __shfl_sync(0xffffffffu, nbase, 0);
                node_off = nbase + x - wcnt;
            }
            if (slot) {
                const uint32_t e = slot_base_s + q;
                uint32_t wf = wcnt & RM_W_MASK;
                if (has) wf |= RM_VALID;
                if (single_pass && has) a.row_key[e] = (uint16_t)(k - (inv_total ? (uint32_t)inv_pre[k] : 0u));
                if (labelled) wf |= RM_LABELLED;
                if (eligible) wf |= RM_ELIGIBLE;
                if (r.monotone) wf |= RM_MONOTONE;
                a.meta_b[e] = make_uint4(node_off, wf, label, labelled ? r.h.hi : 0u);
                if (labelled) {
                    a.hash_lo[e] = r.h.lo;
                    if (eligible) a.meta_a[e] = make_longlong2(r.c8, r.c9);
                }
            }
            if (wcnt && (!LONG || slow)) {
                uint32_t* dst = a.nodes + node_off;
                if (!LONG && r.stashed) {
                    for (uint32_t i = 0; i < wcnt; ++i) dst[i] = stash[i * INGEST_THREADS + tid];
                } else {  // walk longer than the stash: decode it again
                    WalkIter it{b, r.path_pos, r.path_end};
                    int64_t m;
                    for (uint32_t i = 0; i < wcnt; ++i) { it.next(m); dst[i] = (uint32_t)m; }
                }
            }
            __syncwarp();
        }
        valid_prev += n_round - inv_total;
        __syncthreads();
    }
    if (hist_smem) {
        for (uint32_t i = tid; i < HIST_SLOTS; i += INGEST_THREADS) {
            const uint32_t label = hkey[i];
            if (label == LABEL_U) continue;
            unsigned long long* hp = a.hist + 4ull * label;
            const uint32_t* hv = hval + 4u * i;
            atomicAdd(hp + 0, (unsigned long long)hv[0]);
            if (hv[3]) atomicAdd(hp + 1, (unsigned long long)hv[3]);
            if (hv[1]) atomicAdd(hp + 2, (unsigned long long)hv[1]);
            if (hv[2]) atomicAdd(hp + 3, (unsigned long long)hv[2]);
        }
    }
}

// ---- species counts of the records a warp holds (profile.rs:219-232, 264-277): reads, sum of read lengths, #(3 <= mapq <= 60),
// #(mapq == 60).  The lanes of one species add ONCE: the whole warp when it is one species (short reads of a single-species
// sample), otherwise the groups match.any finds.  With hundreds of species a tile's ~124 records are nearly all different species,
// so a per-tile table in shared memory (round 1) merged nothing and cost its initialisation, a CAS per record and a flush of ~4
// global atomics per distinct species; an abundant species still collapses to one set of REDs per warp.
__device__ __forceinline__ void count_species_warp(unsigned long long* hist, bool cnt, uint32_t label, unsigned long long ql, bool lm, bool uq,
                                                   uint32_t lane) {
    const unsigned mm = __ballot_sync(0xffffffffu, cnt);
    if (mm == 0u) return;
    const int leader0 = __ffs(mm) - 1;
    const uint32_t lab0 = __shfl_sync(0xffffffffu, label, leader0);
    const bool uniform = __all_sync(0xffffffffu, !cnt || label == lab0);
    const uint32_t c_lm = (cnt && lm) ? 1u : 0u, c_uq = (cnt && uq) ? 1u : 0u;
    const unsigned long long qv = cnt ? ql : 0ull;
    if (uniform) {
        const uint32_t n1 = __popc(mm);
        const uint32_t n3 = __reduce_add_sync(0xffffffffu, c_lm);
        const uint32_t n4 = __reduce_add_sync(0xffffffffu, c_uq);
        unsigned long long s;
        if (__all_sync(0xffffffffu, qv < (1ull << 26))) {  // 32 x 2^26 fits 32 bits: one REDUX
            s = (unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)qv);
        } else {
            s = qv;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        }
        if ((int)lane == leader0) {
            unsigned long long* hp = hist + 4ull * lab0;
            atomicAdd(hp + 0, (unsigned long long)n1);
            atomicAdd(hp + 1, s);
            if (n3) atomicAdd(hp + 2, (unsigned long long)n3);
            if (n4) atomicAdd(hp + 3, (unsigned long long)n4);
        }
        return;
    }
    const unsigned b_lm = __ballot_sync(0xffffffffu, c_lm != 0u), b_uq = __ballot_sync(0xffffffffu, c_uq != 0u);
    const unsigned peers = __match_any_sync(0xffffffffu, cnt ? label : (0x80000000u | lane));  // (labels are species indices < 2^31; idle lanes stay alone)
    if (!cnt) return;
    unsigned long long s = 0;
    for (unsigned m = peers; m; m &= m - 1u) s += __shfl_sync(peers, qv, __ffs(m) - 1);  // the lanes of a group run the same trip count
    if ((int)lane == __ffs(peers) - 1) {
        unsigned long long* hp = hist + 4ull * label;
        const uint32_t n3 = __popc(b_lm & peers), n4 = __popc(b_uq & peers);
        atomicAdd(hp + 0, (unsigned long long)__popc(peers));
        atomicAdd(hp + 1, s);
        if (n3) atomicAdd(hp + 2, (unsigned long long)n3);
        if (n4) atomicAdd(hp + 3, (unsigned long long)n4);
    }
}

// =====================================================================================
// k_ingest_s: the short-read ingest kernel around a structural index of the tile.
//   A. every thread classifies 16-byte pieces of the staged text (LDS.128, conflict-free): newline and tab flags,
//      byte-ordered, into two shared-memory bitmaps (ptx_fast.cuh: classify16)
//   B. line starts from the newline bitmap (popc + block scan), ordered by line length
//   C. one thread per record: fast_parse (ptx_fast.cuh) finds the columns with ffs on the tab bitmap, converts integers
//      and walk ids four digits per multiply, hashes the id from 4-byte words - no byte is loaded on its own.
//      Records it declines (signs, "\r\n", 10+ digit ids, short rows, lines leaving the window: rare) go through the
//      exact byte parser parse_record, first on the staged text, then on the global copy of the line.
// Same outputs as k_ingest<false>: species counts, record table, CSR walks, labels / tile_info / row_key.
// =====================================================================================
constexpr uint32_t STAGE_PAD = 128;  // '\n' bytes readable behind the staged window (ld8 near its end, sentinel for parse_record)

__device__ __noinline__ void parse_record_exact(const uint8_t* stage, uint32_t p, uint32_t stage_bytes, const uint8_t* gtile, uint32_t glim,
                                                RecParse* out, uint32_t* from_global) {
    RecParse r;
    const uint32_t self = 1u << (threadIdx.x & 31u);
    *from_global = 0u;
    if (!parse_record(stage, p, stage_bytes, r, self, nullptr, 0, 0)) {  // the columns run past the staged window
        parse_record(gtile, p, glim, r, self, nullptr, 0, 0);
        *from_global = 1u;
    }
    *out = r;
}

template <bool CLS_MUL>
__global__ void __launch_bounds__(SHORT_THREADS, PTX_SHORT_MINB) k_ingest_s(const IngestArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
#ifdef PTX_DEBUG_TMA
    long long dbg_t0 = 0;
    const long long dbg_tk = clock64();
#endif
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile_bytes = a.tile_bytes;                 // multiple of 128
    const uint32_t stage_bytes = tile_bytes + a.over_bytes;   // multiple of 128
    const uint32_t bm_words = stage_bytes / 32u;              // bitmap words of the window; a sentinel word of all ones behind them
    uint8_t* stage = smem;                                    // [stage_bytes + STAGE_PAD]
    uint32_t* nlw = reinterpret_cast<uint32_t*>(smem + stage_bytes + STAGE_PAD);             // 16-byte aligned
    uint32_t* tabw = nlw + ((bm_words + 2u + 3u) & ~3u);
    uint32_t* stash = tabw + ((bm_words + 2u + 3u) & ~3u);                                   // [SHORT_STASH_CAP][SHORT_THREADS]
    uint16_t* rec_tmp = reinterpret_cast<uint16_t*>(stash);                                  // [SHORT_REC_CAP] sort scratch, dead before the records are parsed
    uint16_t* rec_start = reinterpret_cast<uint16_t*>(stash + SHORT_STASH_CAP * SHORT_THREADS);  // [SHORT_REC_CAP] line starts of the round
    uint16_t* order = rec_start + SHORT_REC_CAP;                                             // [SHORT_REC_CAP] lines by length
    uint16_t* inv_pre = order + SHORT_REC_CAP;                                               // [SHORT_REC_CAP] invalid line slots before slot k
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t warp_tot[SHORT_THREADS / 32];
    __shared__ uint32_t bin_cnt[64];
    __shared__ uint32_t inv_flag, inv_tot_s, slot_base_s, node_base_s, node_warp[SHORT_THREADS / 32];
    const Words Wd{smem_u32(stage), nullptr}, Wtab{smem_u32(tabw), nullptr}, Wnl{smem_u32(nlw), nullptr};

    const uint64_t t0 = (uint64_t)blockIdx.x * tile_bytes;
    const uint8_t* gtile = a.text + t0;

    // ---- stage the tile: one TMA bulk copy, completion on an mbarrier every thread waits on
    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        inv_flag = 0;
    }
    __syncthreads();
    if (tid == 0) {
#ifdef PTX_DEBUG_TMA
        dbg_t0 = clock64();
#endif
        mbar_expect_tx(&mbar, stage_bytes);
        bulk_g2s_hint(stage, gtile, stage_bytes, &mbar, a.pol_stream);  // the text is read once
        // the CTAs of a wave start together and would all wait for DRAM: fetch the tile of the CTA that takes this one's place into the L2 now
        const uint64_t pf = (uint64_t)blockIdx.x + a.pf_dist;
        if (a.pf_dist && pf < a.n_tiles) {
            const uint64_t off = pf * tile_bytes, lim = a.padded_bytes - off;
            bulk_prefetch_l2(a.text + off, (uint32_t)min((uint64_t)stage_bytes, lim) & ~15u);
        }
    }
    if (tid < STAGE_PAD / 4u) reinterpret_cast<uint32_t*>(stage + stage_bytes)[tid] = 0x0a0a0a0au;
    if (tid < 2u) { nlw[bm_words + tid] = 0xFFFFFFFFu; tabw[bm_words + tid] = 0xFFFFFFFFu; }
    if (tid < 64u) bin_cnt[tid] = 0;
    const uint32_t* sstart = a.ranges.sstart;
    __shared__ uint32_t pv[CLASSIFY_PIVOTS];  // every stride-th range start: the first level of the species search
    const int pv_stride = a.ranges.S > CLASSIFY_PIVOTS ? (a.ranges.S + CLASSIFY_PIVOTS - 1) / CLASSIFY_PIVOTS : 1;
    const int npv = (a.ranges.disjoint && a.ranges.S > 1) ? (a.ranges.S + pv_stride - 1) / pv_stride : 0;
    for (int i = (int)tid; i < npv; i += SHORT_THREADS) pv[i] = sstart[i * pv_stride];  // (visible behind the barrier that follows the structural index)
    const bool single_pass = a.micro_base == nullptr;
    const uint32_t rec_base = single_pass ? 0u : (uint32_t)a.micro_base[(uint64_t)blockIdx.x * (tile_bytes / MICRO)];
    const uint32_t glim = (uint32_t)min((uint64_t)0xFFFF0000ull, a.padded_bytes - t0);
    const RangesView& R = a.ranges;
    mbar_wait(&mbar, 0);
#ifdef PTX_DEBUG_TMA
    if (tid == 0 && (blockIdx.x % 5003u) == 7u) printf("tile %u sm %u: tma issue->ready %lld cycles, kernel start->issue %lld\n", blockIdx.x, (unsigned)__smid(), (long long)(clock64() - dbg_t0), (long long)(dbg_t0 - dbg_tk));
#endif

    // ---- A: structural index of the window (the pad and the sentinel words written above become visible at the barrier behind it)
    for (uint32_t pc = tid; pc < stage_bytes / 16u; pc += SHORT_THREADS) {
        const uint4 q = reinterpret_cast<const uint4*>(stage)[pc];
        uint32_t nl16, tab16;
        if (CLS_MUL) classify16_mul(q.x, q.y, q.z, q.w, nl16, tab16); else classify16(q.x, q.y, q.z, q.w, nl16, tab16);
        reinterpret_cast<uint16_t*>(nlw)[pc] = (uint16_t)nl16;
        reinterpret_cast<uint16_t*>(tabw)[pc] = (uint16_t)tab16;
    }
    __syncthreads();

    // a line belongs to the tile that holds the newline in front of it: line starts q + 1 for newlines at q < tile_bytes,
    // as long as the start lies inside the text (what follows the last line is newline padding)
    const uint32_t nw = tile_bytes / 32u;                               // bitmap words of the tile (a multiple of 4)
    const uint64_t rest = a.n_bytes - t0;                               // text bytes from the tile start (>= 1)
    const uint32_t qmax = rest - 1u < (uint64_t)tile_bytes ? (uint32_t)(rest - 1u) : tile_bytes;  // newlines at q < qmax start a line
    const uint32_t first_slot = (blockIdx.x == 0 && tid == 0) ? 1u : 0u;  // the line at text[0]

    uint32_t n_rec = 0;
    uint32_t valid_prev = 0;
    for (uint32_t round = 0; round == 0 || round < n_rec; round += SHORT_REC_CAP) {
        // ---- B: number the line starts.  Four consecutive bitmap words (one LDS.128) per thread and pass; a tile of up
        // to 16 KB is one pass.  (Recomputed in the rare extra rounds of a tile with more than SHORT_REC_CAP lines.)
        uint32_t idx_base = 0;  // line slots in front of this thread's words, summed over the passes
        for (uint32_t w0 = 0; w0 < nw; w0 += 4u * SHORT_THREADS) {
            const uint32_t wi = w0 + 4u * tid;
            uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
            if (wi < nw) {
                m4 = reinterpret_cast<const uint4*>(nlw)[wi >> 2];
                if ((wi + 4u) * 32u > qmax) {  // last tile only: drop the newlines that start no line
                    uint32_t* mm = &m4.x;
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) {
                        const uint32_t b0 = (wi + j) * 32u;
                        if (b0 + 32u > qmax) mm[j] = b0 >= qmax ? 0u : (mm[j] & ((1u << (qmax - b0)) - 1u));
                    }
                }
            }
            const uint32_t cnt = __popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w) + (w0 == 0u ? first_slot : 0u);
            uint32_t x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            if (w0) __syncthreads();  // the previous pass has read warp_tot
            if (lane == 31u) warp_tot[warp] = x;
            __syncthreads();
            uint32_t idx = idx_base + x - cnt, tot = 0;
#pragma unroll
            for (int w = 0; w < SHORT_THREADS / 32; ++w) {
                const uint32_t t = warp_tot[w];
                if ((uint32_t)w < warp) idx += t;
                tot += t;
            }
            idx_base += tot;
            if (w0 == 0u && first_slot) {
                if (idx >= round && idx < round + SHORT_REC_CAP) rec_start[idx - round] = 0;
                ++idx;
            }
            const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                uint32_t m = mm[j];
                const uint32_t b0 = (wi + j) * 32u + 1u;
                while (m) {
                    const uint32_t bit = __ffs(m) - 1;
                    m &= m - 1u;
                    const uint32_t q = b0 + bit;
                    if (idx >= round && idx < round + SHORT_REC_CAP) {
                        rec_start[idx - round] = (uint16_t)q;
                        if (!valid_first(stage, q)) inv_flag = 1;  // an empty line or an '@' comment: rare
                    }
                    ++idx;
                }
            }
        }
        n_rec = idx_base;
        const uint32_t n_round = min(SHORT_REC_CAP, n_rec - round);
        if (tid == 0 && first_slot && round == 0 && !valid_first(stage, 0)) inv_flag = 1;
        if (single_pass && n_rec > SHORT_REC_CAP) {  // rows are numbered per tile: the host redoes the chunk with the count pass (uniform: every thread knows n_rec)
            if (tid == 0) atomicOr(a.cursors + 3, 1u);
            return;
        }
        __syncthreads();

        // ---- empty lines and '@' comments are line slots but not records (rare): inv_pre[k] = invalid slots before slot k
        uint32_t inv_total = 0;
        if (inv_flag) {
            if (warp == 0) {
                uint32_t base = 0;
                for (uint32_t k0 = 0; k0 < n_round; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    const bool bad = k < n_round && !valid_first(stage, rec_start[k]);
                    const unsigned bm = __ballot_sync(0xffffffffu, bad);
                    if (k < n_round) inv_pre[k] = (uint16_t)(base + __popc(bm & ((1u << lane) - 1u)));
                    base += __popc(bm);
                }
                if (lane == 0) inv_tot_s = base;
            }
            __syncthreads();
            inv_total = inv_tot_s;
            __syncthreads();
            if (tid == 0) inv_flag = 0;  // for the next round
        }

        // ---- lines ordered by length (a proxy for the walk length): counting sort over 64 four-byte bins
        if (a.no_sort) {
            for (uint32_t k = tid; k < n_round; k += SHORT_THREADS) order[k] = (uint16_t)k;
        } else {
            for (uint32_t k = tid; k < n_round; k += SHORT_THREADS) {
                const uint32_t s0 = rec_start[k];
                const uint32_t e0 = (k + 1 < n_round) ? rec_start[k + 1] : s0 + 112u;
                const uint32_t len = e0 - s0;
                const uint32_t bin = len < 64u ? 0u : min((len - 64u) >> 2, 63u);
                rec_tmp[k] = (uint16_t)((bin << 10) | atomicAdd(&bin_cnt[bin], 1u));
            }
            __syncthreads();
            // every warp scans the 64 bin counts for itself (two bins per lane): no barrier around a dedicated scan step
            const uint32_t c0 = bin_cnt[2 * lane], c1 = bin_cnt[2 * lane + 1];
            uint32_t x = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            const uint32_t base_even = x - c0 - c1, base_odd = x - c1;  // first position of bins 2*lane and 2*lane + 1
            for (uint32_t k0 = 0; k0 < n_round; k0 += SHORT_THREADS) {  // uniform trip count: the shuffles need every lane
                const uint32_t k = k0 + tid;
                const uint32_t t = k < n_round ? rec_tmp[k] : 0u, bin = t >> 10;
                const uint32_t be = __shfl_sync(0xffffffffu, base_even, bin >> 1), bo = __shfl_sync(0xffffffffu, base_odd, bin >> 1);
                if (k < n_round) order[((bin & 1u) ? bo : be) + (t & 1023u)] = (uint16_t)k;
            }
        }
        __syncthreads();
        if (tid < 64u) bin_cnt[tid] = 0;  // for the next round (read again only behind further barriers)

        // ---- C: one thread per record
        for (uint32_t k0 = 0; k0 < n_round; k0 += SHORT_THREADS) {
            const uint32_t q = k0 + tid;
            const bool slot = q < n_round;
            const uint32_t k = slot ? order[q] : 0u;
            const uint32_t p = slot ? rec_start[k] : 0u;
            const bool has = slot && valid_first(stage, p);
            const uint32_t pmask = __ballot_sync(0xffffffffu, has);
            // what the rest of the iteration needs to know about the record
            IdHash h;
            h.lo = 0; h.hi = 0;
            uint32_t W = 0, path_pos = 0, path_end = 0;
            int64_t vmin = -1, vmax = -1, c8 = 0, c9 = 0;
            unsigned long long ql = 0;      // read length, 0 if null
            bool lm = false, uq = false;    // 3 <= mapq <= 60, mapq == 60
            bool cols_ok = false;           // path, c7, c8, c9 all non-null (profile.rs:380-399)
            bool monotone = true, stashed = false;
            const uint8_t* b = stage;
            bool fast = false;
            if (has) {
                FastRec f;
                uint32_t e;  // the '\n' that ends the line: in front of the next line start, or (last line of the round) from the bitmap
                if (k + 1u < n_round) e = (uint32_t)rec_start[k + 1u] - 1u;
                else { BitCursor nc; nc.seek(Wnl, p); e = nc.next(); }
                fast = fast_parse(Wd, Wtab, p, e, stage_bytes, f, pmask, stash + tid, SHORT_THREADS, SHORT_STASH_CAP);
                if (fast) {
                    h = f.h;
                    W = f.W;
                    if (W) { vmin = (int64_t)f.vmin; vmax = (int64_t)f.vmax; }
                    c8 = (int64_t)f.c8;
                    c9 = (int64_t)f.c9;
                    ql = (f.nulls & FN_QLEN) ? 0ull : (unsigned long long)f.qlen;
                    lm = !(f.nulls & FN_MAPQ) && f.mapq - 3u <= 57u;
                    uq = lm && f.mapq == 60u;
                    cols_ok = !f.path_null && !(f.nulls & (FN_C7 | FN_C8 | FN_C9));
                    monotone = false;  // not tracked on the fast path: k_apply notices repeats while it walks (cover_record)
                    stashed = f.W <= SHORT_STASH_CAP;  // a longer walk (rare) is decoded again from the staged text when it is written out
                    path_pos = f.path_pos;
                    path_end = f.path_end;
                }
            }
            __syncwarp();
            if (has && !fast) {  // rare: the exact byte parser
                RecParse r;
                uint32_t from_global;
                parse_record_exact(stage, p, stage_bytes, gtile, glim, &r, &from_global);
                if (from_global) b = gtile;
                h = r.h;
                W = r.W;
                if (W) { vmin = r.vmin; vmax = r.vmax; }
                c8 = r.c8;
                c9 = r.c9;
                ql = r.qlen != NULL_I64 ? (unsigned long long)r.qlen : 0ull;
                lm = r.mapq != NULL_I64 && r.mapq >= 3 && r.mapq <= 60;
                uq = lm && r.mapq == 60;
                cols_ok = !r.path_null && r.c7 != NULL_I64 && r.c8 != NULL_I64 && r.c9 != NULL_I64;
                monotone = r.monotone;
                path_pos = r.path_pos;
                path_end = r.path_end;
            }
            __syncwarp();
            uint32_t label = LABEL_U;
            if (has) {
                const uint32_t row = rec_base + valid_prev + k - (inv_total ? (uint32_t)inv_pre[k] : 0u);  // GAF row within the chunk
                if (a.labels_in) {  // strain-only resume: the species column of reads_classification.tsv
                    label = a.labels_in[row];
                    if (label != LABEL_U && W && (vmin < R.start[label] || vmax > R.end[label])) {
                        atomicOr(a.flags + 3, 1u);
                        label = LABEL_U;
                    }
                } else {
                    label = classify_pivots(R, vmin, vmax, sstart, pv, npv, pv_stride);
                }
                if (!single_pass) a.labels[row] = label;
            }
            __syncwarp();
            count_species_warp(a.hist + (size_t)(blockIdx.x % a.hist_copies) * a.hist_stride, has && label != LABEL_U, label, ql, lm, uq, lane);  // species counts (profile.rs:219-232, 264-277)
            // ---- emit the record for k_apply: id hash, label, alignment interval and the walk as CSR node ids
            const bool labelled = has && label != LABEL_U;
            const bool eligible = labelled && cols_ok;
            const uint32_t wcnt = eligible ? W : 0u;
            uint32_t node_off;
            {
                // record-table entries of the round and CSR slots of its walks: ONE 64-bit atomicAdd per tile on the packed cursor
                // {entries, nodes} (the cursor is a single hot address of the whole grid: per-warp atomics on it made every warp wait
                // for a serialised L2 round trip).  Block scan of the walk lengths: warp scans + the warp totals in shared memory.
                uint32_t x = wcnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                if (lane == 31u) node_warp[warp] = x;
                __syncthreads();
                if (tid == 0) {
                    uint32_t tot = 0;
#pragma unroll
                    for (int w = 0; w < SHORT_THREADS / 32; ++w) tot += node_warp[w];
                    const uint32_t n_ent = k0 == 0u ? n_round : 0u;  // the first pass over the round takes its entries
                    uint32_t nb = 0, sb = slot_base_s;
                    if (tot | n_ent) {
                        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(a.cursors), ((unsigned long long)tot << 32) | n_ent);
                        nb = (uint32_t)(old >> 32);
                        if (n_ent) sb = (uint32_t)old;
                    }
                    // the table / the CSR buffer was sized from an estimate: the host redoes the chunk with exact sizes
                    if ((single_pass && n_ent && sb + n_ent > a.slots_cap) || nb + tot > a.nodes_cap) {
                        atomicOr(a.cursors + 3, 1u);
                        nb = 0xFFFFFFFFu;
                    } else if (n_ent && single_pass) {
                        a.tile_info[blockIdx.x] = make_uint4(sb, n_round, n_round - inv_total, 0u);
                        atomicAdd(a.cursors + 2, n_round - inv_total);
                    }
                    slot_base_s = sb;
                    node_base_s = nb;
                }
                __syncthreads();
                if (node_base_s == 0xFFFFFFFFu) return;  // uniform; nothing of an abandoned chunk counts (k_apply, k_hist_merge test cursors[3])
                uint32_t before = node_base_s + x - wcnt;
#pragma unroll
                for (int w = 0; w < SHORT_THREADS / 32; ++w)
                    if ((uint32_t)w < warp) before += node_warp[w];
                node_off = before;
            }
            if (slot) {
                const uint32_t e = slot_base_s + q;
                uint32_t wf = wcnt & RM_W_MASK;
                if (has) wf |= RM_VALID;
                // the record table streams out (k_apply reads it once): evict-first stores
                if (single_pass && has) __stcs(a.row_key + e, (uint16_t)(k - (inv_total ? (uint32_t)inv_pre[k] : 0u)));
                if (labelled) wf |= RM_LABELLED;
                if (eligible) wf |= RM_ELIGIBLE;
                if (monotone) wf |= RM_MONOTONE;
                __stcs(a.meta_b + e, make_uint4(node_off, wf, label, labelled ? h.hi : 0u));
                if (labelled) {
                    __stcs(a.hash_lo + e, (unsigned long long)h.lo);
                    if (eligible) __stcs(a.meta_a + e, make_longlong2(c8, c9));
                }
            }
            if (wcnt) {
                uint32_t* dst = a.nodes + node_off;
                if (stashed) {
                    for (uint32_t i = 0; i < wcnt; ++i) __stcs(dst + i, stash[i * SHORT_THREADS + tid]);
                } else {  // exact parser: decode the walk again
                    WalkIter it{b, path_pos, path_end};
                    int64_t m;
                    for (uint32_t i = 0; i < wcnt; ++i) { it.next(m); dst[i] = (uint32_t)m; }
                }
            }
            __syncwarp();
        }
        valid_prev += n_round - inv_total;
        __syncthreads();
    }
}

// =====================================================================================
// k_ingest_l: the long-read ingest kernel (HiFi / ONT: lines of hundreds of bytes, walks of tens to hundreds of nodes).
// Same structural index as k_ingest_s plus a non-digit bitmap; the walk column is decoded NODE-parallel:
//   C1  one thread per line: columns 1-5 and 7-12 (fast_cols); the extent [p6, e6] of its walk column goes to shared memory
//   C2  every thread takes four words of the bitmaps and finds the lines they meet (binary search over the line starts): a walk
//       id ENDS where a non-digit byte inside a walk column follows a digit; popc + block scan number the ids of the tile in
//       file order (= their CSR slots, one atomicAdd per tile); each end is converted by the thread that owns its bit (run
//       start from the non-digit bitmap, SWAR digits) and stored at its slot - a thread's ids are consecutive slots; the
//       smallest / largest id of each line accumulate in shared memory (atomicMin / atomicMax per thread and line)
//   C3  one thread per line again: CSR offset and node count from the ranks of p6 and e6, label, species counts,
//       record-table entry
// Lines the word-wide path declines (fast_cols) or whose walk has a 10+ digit id go through the exact byte parser.
// =====================================================================================
constexpr uint32_t LONG_WPT = (MAX_TILE + OVER) / 32u / LONG_THREADS + 1u;  // bitmap words a thread of k_ingest_l takes

__global__ void __launch_bounds__(LONG_THREADS, PTX_LONG_MINB) k_ingest_l(const IngestArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile_bytes = a.tile_bytes;                 // multiple of 4096
    const uint32_t stage_bytes = tile_bytes + a.over_bytes;   // multiple of 128
    const uint32_t bm_words = stage_bytes / 32u;
    const uint32_t bm_alloc = (bm_words + 2u + 3u) & ~3u;
    uint8_t* stage = smem;                                    // [stage_bytes + STAGE_PAD]
    uint32_t* nlw = reinterpret_cast<uint32_t*>(smem + stage_bytes + STAGE_PAD);
    uint32_t* tabw = nlw + bm_alloc;
    uint32_t* ndw = tabw + bm_alloc;    // non-digit bytes
    uint32_t* ew = ndw + bm_alloc;      // the ends of the walk ids of the group's lines
    uint16_t* epre = reinterpret_cast<uint16_t*>(ew + bm_alloc);  // ids that end in front of each bitmap word (exclusive prefix over the tile: < 2^16, an id takes two bytes)
    uint32_t* lmin = reinterpret_cast<uint32_t*>(epre + bm_alloc);                                                        // [LONG_THREADS] smallest / largest walk id of each line of the group
    uint32_t* lmax = lmin + LONG_THREADS;                                                  // [LONG_THREADS]
    uint16_t* lp6 = reinterpret_cast<uint16_t*>(lmax + LONG_THREADS);                      // [LONG_THREADS] first byte of the line's walk column (line start if it has none)
    uint16_t* lend = lp6 + LONG_THREADS;                                                   // [LONG_THREADS] one past its closing tab (<= lp6: no column)
    uint16_t* rec_start = lend + LONG_THREADS;                                             // [LONG_REC_CAP]
    uint16_t* inv_pre = rec_start + LONG_REC_CAP;                                                 // [LONG_REC_CAP]
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t warp_tot[LONG_THREADS / 32];
    __shared__ uint32_t inv_flag, inv_tot_s, slot_base_s, node_base_s, slow_bits[LONG_THREADS / 32];
    const Words Wd{smem_u32(stage), nullptr}, Wtab{smem_u32(tabw), nullptr}, Wnl{smem_u32(nlw), nullptr};

    const uint64_t t0 = (uint64_t)blockIdx.x * tile_bytes;
    const uint8_t* gtile = a.text + t0;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        inv_flag = 0;
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&mbar, stage_bytes);
        bulk_g2s_hint(stage, gtile, stage_bytes, &mbar, a.pol_stream);  // the text is read once
        // the CTAs of a wave start together and would all wait for DRAM: fetch the tile of the CTA that takes this one's place into the L2 now
        const uint64_t pf = (uint64_t)blockIdx.x + a.pf_dist;
        if (a.pf_dist && pf < a.n_tiles) {
            const uint64_t off = pf * tile_bytes, lim = a.padded_bytes - off;
            bulk_prefetch_l2(a.text + off, (uint32_t)min((uint64_t)stage_bytes, lim) & ~15u);
        }
    }
    if (tid < STAGE_PAD / 4u) reinterpret_cast<uint32_t*>(stage + stage_bytes)[tid] = 0x0a0a0a0au;
    if (tid < 2u) { nlw[bm_words + tid] = 0xFFFFFFFFu; tabw[bm_words + tid] = 0xFFFFFFFFu; ndw[bm_words + tid] = 0xFFFFFFFFu; ew[bm_words + tid] = 0u; epre[bm_words + tid] = 0; }
    const uint32_t* sstart = a.ranges.sstart;
    __shared__ uint32_t pv[CLASSIFY_PIVOTS];  // every stride-th range start: the first level of the species search
    const int pv_stride = a.ranges.S > CLASSIFY_PIVOTS ? (a.ranges.S + CLASSIFY_PIVOTS - 1) / CLASSIFY_PIVOTS : 1;
    const int npv = (a.ranges.disjoint && a.ranges.S > 1) ? (a.ranges.S + pv_stride - 1) / pv_stride : 0;
    for (int i = (int)tid; i < npv; i += LONG_THREADS) pv[i] = sstart[i * pv_stride];  // (visible behind the barrier that follows the structural index)
    const bool single_pass = a.micro_base == nullptr;
    const uint32_t rec_base = single_pass ? 0u : (uint32_t)a.micro_base[(uint64_t)blockIdx.x * (tile_bytes / MICRO)];
    const uint32_t glim = (uint32_t)min((uint64_t)0xFFFF0000ull, a.padded_bytes - t0);
    const RangesView& R = a.ranges;
    mbar_wait(&mbar, 0);

    // ---- A: structural index of the window: newline, tab and non-digit flags
    for (uint32_t pc = tid; pc < stage_bytes / 16u; pc += LONG_THREADS) {
        const uint4 q = reinterpret_cast<const uint4*>(stage)[pc];
        uint32_t nl16, tab16;
        classify16(q.x, q.y, q.z, q.w, nl16, tab16);
        reinterpret_cast<uint16_t*>(nlw)[pc] = (uint16_t)nl16;
        reinterpret_cast<uint16_t*>(tabw)[pc] = (uint16_t)tab16;
        reinterpret_cast<uint16_t*>(ndw)[pc] = (uint16_t)nondigit16(q.x, q.y, q.z, q.w);
    }
    __syncthreads();

    const uint32_t nw = tile_bytes / 32u;
    const uint64_t rest = a.n_bytes - t0;
    const uint32_t qmax = rest - 1u < (uint64_t)tile_bytes ? (uint32_t)(rest - 1u) : tile_bytes;
    const uint32_t first_slot = (blockIdx.x == 0 && tid == 0) ? 1u : 0u;

    uint32_t n_rec = 0;
    uint32_t valid_prev = 0;
    for (uint32_t round = 0; round == 0 || round < n_rec; round += LONG_REC_CAP) {
        // ---- B: number the line starts (as k_ingest_s)
        uint32_t idx_base = 0;
        for (uint32_t w0 = 0; w0 < nw; w0 += 4u * LONG_THREADS) {
            const uint32_t wi = w0 + 4u * tid;
            uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
            if (wi < nw) {
                m4 = reinterpret_cast<const uint4*>(nlw)[wi >> 2];
                if ((wi + 4u) * 32u > qmax) {
                    uint32_t* mm = &m4.x;
#pragma unroll
                    for (uint32_t j = 0; j < 4u; ++j) {
                        const uint32_t b0 = (wi + j) * 32u;
                        if (b0 + 32u > qmax) mm[j] = b0 >= qmax ? 0u : (mm[j] & ((1u << (qmax - b0)) - 1u));
                    }
                }
            }
            const uint32_t cnt = __popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w) + (w0 == 0u ? first_slot : 0u);
            uint32_t x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            if (w0) __syncthreads();
            if (lane == 31u) warp_tot[warp] = x;
            __syncthreads();
            uint32_t idx = idx_base + x - cnt, tot = 0;
#pragma unroll
            for (int w = 0; w < LONG_THREADS / 32; ++w) {
                const uint32_t t = warp_tot[w];
                if ((uint32_t)w < warp) idx += t;
                tot += t;
            }
            idx_base += tot;
            if (w0 == 0u && first_slot) {
                if (idx >= round && idx < round + LONG_REC_CAP) rec_start[idx - round] = 0;
                ++idx;
            }
            const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j) {
                uint32_t m = mm[j];
                const uint32_t b0 = (wi + j) * 32u + 1u;
                while (m) {
                    const uint32_t bit = __ffs(m) - 1;
                    m &= m - 1u;
                    const uint32_t q = b0 + bit;
                    if (idx >= round && idx < round + LONG_REC_CAP) {
                        rec_start[idx - round] = (uint16_t)q;
                        if (!valid_first(stage, q)) inv_flag = 1;
                    }
                    ++idx;
                }
            }
        }
        n_rec = idx_base;
        const uint32_t n_round = min(LONG_REC_CAP, n_rec - round);
        if (tid == 0 && first_slot && round == 0 && !valid_first(stage, 0)) inv_flag = 1;
        if (single_pass && n_rec > LONG_REC_CAP) {  // rows are numbered per tile: the host redoes the chunk with the count pass (uniform)
            if (tid == 0) atomicOr(a.cursors + 3, 1u);
            return;
        }
        __syncthreads();
        uint32_t inv_total = 0;
        if (inv_flag) {
            if (warp == 0) {
                uint32_t base = 0;
                for (uint32_t k0 = 0; k0 < n_round; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    const bool bad = k < n_round && !valid_first(stage, rec_start[k]);
                    const unsigned bm = __ballot_sync(0xffffffffu, bad);
                    if (k < n_round) inv_pre[k] = (uint16_t)(base + __popc(bm & ((1u << lane) - 1u)));
                    base += __popc(bm);
                }
                if (lane == 0) inv_tot_s = base;
            }
            __syncthreads();
            inv_total = inv_tot_s;
            __syncthreads();
            if (tid == 0) inv_flag = 0;
        }

        for (uint32_t g0 = 0; g0 < n_round; g0 += LONG_THREADS) {  // groups of one line per thread, in file order
            const uint32_t ng = min((uint32_t)LONG_THREADS, n_round - g0);  // lines of the group
            // ---- C1: the scalar columns of the line; the extent of its walk column goes to lp6 / lend
            if (tid < LONG_THREADS / 32) slow_bits[tid] = 0u;
            const uint32_t k = g0 + tid;
            const bool slot = k < n_round;
            const uint32_t p = slot ? rec_start[k] : 0u;
            const bool has = slot && valid_first(stage, p);
            const uint32_t pmask = __ballot_sync(0xffffffffu, has);
            FastRec f;
            f.nulls = 0; f.W = 0; f.path_pos = f.path_end = 0; f.path_null = true; f.h.lo = 0; f.h.hi = 0;
            f.qlen = f.c7 = f.c8 = f.c9 = f.mapq = 0;
            bool walk = false;  // the word-wide path reads this line: its walk column takes part in C2
            if (has) {
                uint32_t e;
                if (k + 1u < n_round) e = (uint32_t)rec_start[k + 1u] - 1u;
                else { BitCursor nc; nc.seek(Wnl, p); e = nc.next(); }
                uint32_t slowb = 0;
                walk = fast_cols(Wd, Wtab, p, e, stage_bytes, f, pmask, slowb) && slowb == 0u;
            }
            if (slot) {
                lp6[tid] = (uint16_t)(walk ? f.path_pos : p);
                lend[tid] = (uint16_t)(walk ? f.path_end + 1u : p);  // [p6, e6]: the closing tab ends the last id
                lmin[tid] = 0xFFFFFFFFu;
                lmax[tid] = 0u;
            }
            __syncthreads();
            // ---- C2: id ends = non-digit byte of a walk column behind a digit; numbered over the tile (= CSR slots); converted by the
            // thread that owns the bitmap word.  The lines a word meets come from a binary search over the group's line starts.
            {
                // a thread takes `wpt` consecutive bitmap words (at most LONG_WPT: 34 KB of window / 32 / 256 threads)
                const uint32_t wpt = (bm_words + LONG_THREADS - 1u) / LONG_THREADS;
                const uint32_t wi = wpt * tid, wn = wi < bm_words ? min(wpt, bm_words - wi) : 0u;
                uint32_t E[LONG_WPT];
#pragma unroll
                for (uint32_t j = 0; j < LONG_WPT; ++j) E[j] = 0u;
                uint32_t kk0 = 0;  // last line of the group that starts at or before this thread's first word (0 if none)
                if (wn) {
                    const uint32_t wb0 = wi * 32u;
                    uint32_t lo = 0, hi = ng;
                    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint32_t)rec_start[g0 + mid] <= wb0) lo = mid + 1u; else hi = mid; }
                    kk0 = lo ? lo - 1u : 0u;
                    uint32_t kk = kk0;
                    uint32_t carry = wi ? (ndw[wi - 1u] >> 31) : 1u;  // class of the byte in front of the word
#pragma unroll
                    for (uint32_t j = 0; j < LONG_WPT; ++j) {
                        if (j < wn) {
                            const uint32_t nd = ndw[wi + j], wb = (wi + j) * 32u;
                            uint32_t col = 0;  // bytes of this word inside a walk column
                            while (kk < ng) {
                                const uint32_t b0 = lp6[kk], b1 = lend[kk];
                                if (b0 > wb + 31u) break;
                                if (b1 > b0 && b1 > wb) {
                                    const uint32_t lo = b0 > wb ? b0 - wb : 0u, hi = b1 - 1u < wb + 31u ? b1 - 1u - wb : 31u;
                                    col |= (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31u - hi));
                                    if (b1 > wb + 32u) break;  // the column goes on in the next word
                                }
                                ++kk;
                            }
                            E[j] = nd & col & ~((nd << 1) | carry);
                            carry = nd >> 31;
                        }
                    }
                }
                uint32_t cnt = 0;
#pragma unroll
                for (uint32_t j = 0; j < LONG_WPT; ++j) cnt += __popc(E[j]);
                uint32_t x = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                if (lane == 31u) warp_tot[warp] = x;
                __syncthreads();
                uint32_t before = x - cnt, tot = 0;
#pragma unroll
                for (int w = 0; w < LONG_THREADS / 32; ++w) {
                    const uint32_t t = warp_tot[w];
                    if ((uint32_t)w < warp) before += t;
                    tot += t;
                }
                {   // ranks of the ends for C3 (W and CSR offset of a line from the ranks of p6 and e6 + 1)
                    uint32_t c = before;
#pragma unroll
                    for (uint32_t j = 0; j < LONG_WPT; ++j)
                        if (j < wn) { ew[wi + j] = E[j]; epre[wi + j] = (uint16_t)c; c += __popc(E[j]); }
                }
                if (tid == 0) {  // the tile's CSR slots and (first group) the record-table entries of the round: one atomicAdd on the packed cursor
                    const uint32_t n_ent = g0 == 0u ? n_round : 0u;
                    uint32_t nb = 0, sb = slot_base_s;
                    if (tot | n_ent) {
                        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(a.cursors), ((unsigned long long)tot << 32) | n_ent);
                        nb = (uint32_t)(old >> 32);
                        if (n_ent) sb = (uint32_t)old;
                    }
                    if (single_pass && n_ent && sb + n_ent > a.slots_cap) {  // the table was sized from an estimate: the host redoes the chunk
                        atomicOr(a.cursors + 3, 1u);
                        nb = 0xFFFFFFFFu;
                    } else if (n_ent && single_pass) {
                        a.tile_info[blockIdx.x] = make_uint4(sb, n_round, n_round - inv_total, 0u);
                        atomicAdd(a.cursors + 2, n_round - inv_total);
                    }
                    slot_base_s = sb;
                    node_base_s = nb;
                    epre[bm_words] = (uint16_t)tot;
                }
                __syncthreads();
                if (node_base_s == 0xFFFFFFFFu) return;  // uniform
                // sweep: every thread converts the ids that end in its words and stores them at their CSR slots (a thread's ids are
                // consecutive slots: its stores fill whole sectors); min / max per line through shared-memory atomics
                if (cnt) {
                    uint32_t ord = node_base_s + before;
                    uint32_t kk = kk0, cur = 0xFFFFFFFFu, mn = 0xFFFFFFFFu, mx = 0u;
                    bool pending = true;  // a line may have ended since the line of an id was last looked up
#pragma unroll
                    for (uint32_t j = 0; j < LONG_WPT; ++j) {
                        uint32_t m = E[j];
                        const uint32_t nlb = j < wn ? nlw[wi + j] : 0u;
                        if (m) {
                            const uint32_t nd = ndw[wi + j], wb = (wi + j) * 32u;
                            const bool line_may_change = nlb != 0u;
                            bool track = pending || line_may_change;  // look the line up for the first id behind a newline; for every id of a word that holds one
                            pending = false;
                            while (m) {
                                const uint32_t bit = __ffs(m) - 1;
                                m &= m - 1u;
                                const uint32_t q = wb + bit;  // the non-digit byte behind the id
                                if (track) {
                                    while (kk + 1u < ng && (uint32_t)lp6[kk + 1u] <= q) ++kk;  // the line whose column holds q
                                    if (kk != cur) {
                                        if (cur != 0xFFFFFFFFu) { atomicMin(lmin + cur, mn); atomicMax(lmax + cur, mx); }
                                        cur = kk; mn = 0xFFFFFFFFu; mx = 0u;
                                    }
                                    track = line_may_change;
                                }
                                // the id starts behind the previous non-digit byte (the tab in front of the column at the latest)
                                uint32_t below = nd & ((1u << bit) - 1u), wq = wi + j;
                                while (below == 0u) below = ndw[--wq];
                                const uint32_t astart = wq * 32u + (31u - (uint32_t)__clz(below)) + 1u;
                                const uint32_t n = q - astart;
                                uint32_t v = 0;
                                if (n <= 9u) v = fast_node(Wd, astart, n);
                                else atomicOr(&slow_bits[kk >> 5], 1u << (kk & 31u));  // a 10+ digit id: the line goes through the exact parser
                                mn = v < mn ? v : mn;
                                mx = v > mx ? v : mx;
                                __stcs(a.nodes + ord, v);
                                ++ord;
                            }
                        }
                        if (nlb) pending = true;
                    }
                    if (cur != 0xFFFFFFFFu) { atomicMin(lmin + cur, mn); atomicMax(lmax + cur, mx); }
                }
            }
            __syncthreads();  // min / max and slow flags of the group's lines are complete
            // ---- C3: per line: CSR slice, min / max, label, counts, record-table entry
            IdHash h = f.h;
            uint32_t W = 0, node_off = 0, w_res = 0, path_pos = f.path_pos, path_end = f.path_end;
            int64_t vmin = -1, vmax = -1, c8 = (int64_t)f.c8, c9 = (int64_t)f.c9;
            unsigned long long ql = (f.nulls & FN_QLEN) ? 0ull : (unsigned long long)f.qlen;
            bool lm = !(f.nulls & FN_MAPQ) && f.mapq - 3u <= 57u, uq = false;
            uq = lm && f.mapq == 60u;
            bool cols_ok = !f.path_null && !(f.nulls & (FN_C7 | FN_C8 | FN_C9));
            bool monotone = false;  // not tracked: k_apply notices repeats while it walks (cover_record)
            const uint8_t* b = stage;
            bool exact = has && !walk;
            if (has && walk) {
                const uint32_t r0 = epre[path_pos >> 5] + __popc(ew[path_pos >> 5] & ((1u << (path_pos & 31u)) - 1u));
                const uint32_t pe1 = path_end + 1u;
                const uint32_t r1 = epre[pe1 >> 5] + __popc(ew[pe1 >> 5] & ((1u << (pe1 & 31u)) - 1u));
                W = r1 - r0;
                w_res = W;
                node_off = node_base_s + r0;
                if ((slow_bits[tid >> 5] >> (tid & 31u)) & 1u) exact = true;
                else if (W) {
                    vmin = (int64_t)lmin[tid];
                    vmax = (int64_t)lmax[tid];
                }
            }
            __syncwarp();
            if (exact) {  // rare: the exact byte parser
                RecParse r;
                uint32_t from_global;
                parse_record_exact(stage, p, stage_bytes, gtile, glim, &r, &from_global);
                if (from_global) b = gtile;
                h = r.h;
                W = r.W;
                vmin = vmax = -1;
                if (W) { vmin = r.vmin; vmax = r.vmax; }
                c8 = r.c8;
                c9 = r.c9;
                ql = r.qlen != NULL_I64 ? (unsigned long long)r.qlen : 0ull;
                lm = r.mapq != NULL_I64 && r.mapq >= 3 && r.mapq <= 60;
                uq = lm && r.mapq == 60;
                cols_ok = !r.path_null && r.c7 != NULL_I64 && r.c8 != NULL_I64 && r.c9 != NULL_I64;
                monotone = r.monotone;
                path_pos = r.path_pos;
                path_end = r.path_end;
            }
            __syncwarp();
            uint32_t label = LABEL_U;
            if (has) {
                const uint32_t row = rec_base + valid_prev + k - (inv_total ? (uint32_t)inv_pre[k] : 0u);
                if (a.labels_in) {
                    label = a.labels_in[row];
                    if (label != LABEL_U && W && (vmin < R.start[label] || vmax > R.end[label])) {
                        atomicOr(a.flags + 3, 1u);
                        label = LABEL_U;
                    }
                } else {
                    label = classify_pivots(R, vmin, vmax, sstart, pv, npv, pv_stride);
                }
                if (!single_pass) a.labels[row] = label;
            }
            __syncwarp();
            count_species_warp(a.hist + (size_t)(blockIdx.x % a.hist_copies) * a.hist_stride, has && label != LABEL_U, label, ql, lm, uq, lane);  // species counts (profile.rs:219-232, 264-277)
            // ---- the record for k_apply
            const bool labelled = has && label != LABEL_U;
            const bool eligible = labelled && cols_ok;
            const uint32_t wcnt = eligible ? W : 0u;
            // the exact parser finds at most the ids the word-wide pass counted (it drops runs of more than 18 digits): a line that
            // was converted reuses its slots; only a line that never took part in C2 takes new ones - the slots of a chunk stay
            // within text bytes / 2 (chunk_nodes_ensure)
            if (exact && wcnt > w_res) node_off = atomicAdd(a.cursors + 1, wcnt);
            if (slot) {
                const uint32_t e = slot_base_s + k;
                uint32_t wf = wcnt & RM_W_MASK;
                if (has) wf |= RM_VALID;
                // the record table streams out (k_apply reads it once): evict-first stores
                if (single_pass && has) __stcs(a.row_key + e, (uint16_t)(k - (inv_total ? (uint32_t)inv_pre[k] : 0u)));
                if (labelled) wf |= RM_LABELLED;
                if (eligible) wf |= RM_ELIGIBLE;
                if (monotone) wf |= RM_MONOTONE;
                __stcs(a.meta_b + e, make_uint4(node_off, wf, label, labelled ? h.hi : 0u));
                if (labelled) {
                    __stcs(a.hash_lo + e, (unsigned long long)h.lo);
                    if (eligible) __stcs(a.meta_a + e, make_longlong2(c8, c9));
                }
            }
            if (exact && wcnt) {
                uint32_t* dst = a.nodes + node_off;
                WalkIter it{b, path_pos, path_end};
                int64_t m;
                for (uint32_t i = 0; i < wcnt; ++i) { it.next(m); dst[i] = (uint32_t)m; }
            }
            __syncthreads();  // `ew`, `slow_bits`, the per-line arrays are reused by the next group
        }
        valid_prev += n_round - inv_total;
        __syncthreads();
    }
}

// =====================================================================================
// k_apply<MODE>: one thread per record-table entry written by k_ingest.  MODE_CLASSIFY: read-id set insert
// (profile.rs:361-437).  MODE_COVER: node coverage / trio accumulation from the CSR walk (profile.rs:787-919),
// MODE_KEEPMASK: skipping reads whose id group is DS_MIXED.  Small register state, no text: runs at high occupancy,
// and it is also the replay pass (mixed id groups, graphs committed after the ingest) - no text is re-read.
// =====================================================================================
// Block-level append of 16-byte entries to one of n_dest arrays (n_dest <= BOX_STAGE_RANKS): the block's entries are
// grouped by destination in shared memory, one atomicAdd per destination reserves their slots, and the copy-out
// writes consecutive entries from consecutive threads (512-byte bursts per warp).  Every thread of the block calls
// it (dest = 0xFFFFFFFF: nothing to append).  ptr_of(d) = base of array d; an entry beyond `cap` is dropped and reported.
template <class PtrOf, class Dropped>
__device__ __forceinline__ void block_append(uint32_t dest, const ulonglong2& ent, uint32_t n_dest, unsigned long long* cursors, uint64_t cap,
                                             PtrOf ptr_of, Dropped dropped) {
    __shared__ ulonglong2 s_ent[256];
    __shared__ uint32_t s_cnt[BOX_STAGE_RANKS], s_off[BOX_STAGE_RANKS + 1];
    __shared__ unsigned long long s_base[BOX_STAGE_RANKS];
    const uint32_t tid = threadIdx.x;
    if (tid < BOX_STAGE_RANKS) s_cnt[tid] = 0;
    __syncthreads();
    uint32_t my_pos = 0;
    if (dest != 0xFFFFFFFFu) my_pos = atomicAdd(&s_cnt[dest], 1u);
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t d = 0; d < n_dest; ++d) { s_off[d] = run; run += s_cnt[d]; }
        s_off[n_dest] = run;
    }
    if (tid < n_dest && s_cnt[tid]) s_base[tid] = atomicAdd(cursors + tid, (unsigned long long)s_cnt[tid]);
    __syncthreads();
    if (dest != 0xFFFFFFFFu) s_ent[s_off[dest] + my_pos] = ent;
    __syncthreads();
    if (tid < s_off[n_dest]) {
        uint32_t d = 0;
        while (tid >= s_off[d + 1]) ++d;  // at most n_dest steps
        const unsigned long long pos = s_base[d] + (tid - s_off[d]);
        if (pos < cap) ptr_of(d)[pos] = s_ent[tid];
        else dropped();
    }
}

template <int MODE, int VAR = 0>
__global__ void __launch_bounds__(256, PTX_APPLY_MINB) k_apply(const IngestArgs a, uint32_t n_entries) {  // 32 registers, 64 warps/SM: the random id-set and node accesses want every warp they can get (0.947 -> 0.906 ms)
    if (n_entries == ENTRIES_FROM_DEVICE) {  // single-pass ingest: the host does not know the entry count yet
        if (a.cursors[3]) return;            // the estimate was too small: nothing of this chunk counts, it is redone
        n_entries = a.cursors[0];
    }
    __shared__ uint32_t sc_tag[VAR == 2 ? SC_SLOTS : 1];
    __shared__ unsigned long long sc_sum[VAR == 2 ? SC_SLOTS : 1];
    if constexpr (VAR == 2) {
        for (uint32_t i = threadIdx.x; i < SC_SLOTS; i += blockDim.x) { sc_tag[i] = 0xFFFFFFFFu; sc_sum[i] = 0ull; }
        __syncthreads();
    }
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool has = e < n_entries;
    uint4 mb = make_uint4(0u, 0u, LABEL_U, 0u);
    if (has) mb = __ldcs(a.meta_b + e);  // the record table is read once: evict-first
    const bool labelled = has && (mb.y & RM_LABELLED);
    const bool eligible = has && (mb.y & RM_ELIGIBLE);
    const uint32_t label = mb.z;
    IdHash h;
    h.lo = 0;
    h.hi = mb.w;
    if (labelled && (MODE & (MODE_CLASSIFY | MODE_KEEPMASK | MODE_REBOX))) h.lo = __ldcs(a.hash_lo + e);
    if (MODE & (MODE_CLASSIFY | MODE_REBOX)) {
        if (a.box_ptr == nullptr) {
            if ((MODE & MODE_CLASSIFY) && labelled) ds_insert(a.ds, a.ds_shift, a.ds_mask, a.ds_epoch, h, eligible, label, a.flags, a.pol_ds);
        } else {
            // multi-GPU: a read id is kept only by the rank that owns its hash.  Own ids go into the local set; the
            // others are appended to the owner's outbox as {hash, state} (one atomicAdd per distinct owner per warp)
            // and merged there by k_ds_merge_boxes when ptx_finalize exchanges the boxes.
            const ulonglong2 ent = make_ulonglong2(h.lo, ((uint64_t)h.hi << 32) | (eligible ? label : DS_NONE));
            const uint32_t owner = labelled ? ds_owner(ent, a.n_ranks) : 0xFFFFFFFFu;
            const bool mine = labelled && owner == a.rank;
            if ((MODE & MODE_CLASSIFY) && mine) ds_insert(a.ds, a.ds_shift, a.ds_mask, a.ds_epoch, h, eligible, label, a.flags, a.pol_ds);
            const uint32_t dest = (labelled && !mine) ? owner : 0xFFFFFFFFu;
            if (a.n_ranks <= BOX_STAGE_RANKS) {
                block_append(dest, ent, a.n_ranks, a.out_cursor, a.box_cap, [&](uint32_t d) { return a.box_ptr[d]; },
                             [&]() { atomicOr(a.out_cursor + a.n_ranks, 1ull); });  // dropped: sticky marker, all-gathered with the cursors
            } else {
                const unsigned peers = __match_any_sync(0xffffffffu, dest);
                const int leader = __ffs(peers) - 1;
                const uint32_t lane = threadIdx.x & 31u;
                unsigned long long base = 0;
                if (dest != 0xFFFFFFFFu && (int)lane == leader) base = atomicAdd(a.out_cursor + dest, (unsigned long long)__popc(peers));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (dest != 0xFFFFFFFFu) {
                    const unsigned long long pos = base + __popc(peers & ((1u << lane) - 1u));
                    if (pos < a.box_cap) a.box_ptr[dest][pos] = ent;
                    else atomicOr(a.out_cursor + a.n_ranks, 1ull);
                }
            }
        }
    }
    if (MODE & MODE_COVER) {
        const RangesView& R = a.ranges;
        int64_t nb = -1;
        bool keep = false;
        if (eligible) {
            nb = R.node_base[label];
            keep = nb >= 0;
            if ((MODE & MODE_KEEPMASK) && keep) keep = ds_lookup(a.ds, a.ds_shift, a.ds_mask, a.ds_epoch, h, a.pol_ds) != DS_MIXED;  // :415-416
        }
        const uint32_t cmask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const longlong2 se = __ldcs(a.meta_a + e);
            RecParse r;
            r.W = mb.y & RM_W_MASK;
            r.c8 = se.x;
            r.c9 = se.y;
            r.monotone = (mb.y & RM_MONOTONE) != 0;
            r.stashed = true;
            r.path_pos = r.path_end = 0;
            if constexpr (VAR == 1) {
                WarpAggSink sink{{a}};
                cover_record(nullptr, r, label, R.start[label], nb, sink, cmask, a.nodes + mb.x, 1u);
            } else if constexpr (VAR == 2) {
                SmemSink sink{{a}, sc_tag, sc_sum};
                cover_record(nullptr, r, label, R.start[label], nb, sink, cmask, a.nodes + mb.x, 1u);
            } else if constexpr (VAR == 3) {
                PairSink sink{{a}, mb.x};
                cover_record(nullptr, r, label, R.start[label], nb, sink, cmask, a.nodes + mb.x, 1u);
                for (uint32_t q = sink.pos; q < mb.x + r.W; ++q) a.pair_key[q] = 0xFFFFFFFFu;  // later visits of a node, skipped reads
            } else {
                DevSink sink{a};
                cover_record(nullptr, r, label, R.start[label], nb, sink, cmask, a.nodes + mb.x, 1u);
            }
        } else if (VAR == 3 && eligible) {
            for (uint32_t q = mb.x; q < mb.x + (mb.y & RM_W_MASK); ++q) a.pair_key[q] = 0xFFFFFFFFu;  // a read dropped by the keep mask / without a graph
        }
    }
    if constexpr (VAR == 2) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < SC_SLOTS; i += blockDim.x)
            if (sc_tag[i] != 0xFFFFFFFFu) atomicAdd(a.bases + sc_tag[i], sc_sum[i]);
    }
}

// single-pass ingest: the species counts of a chunk go to a chunk-local buffer and are added to the totals only
// if the chunk was not abandoned (cursors[3])
__global__ void __launch_bounds__(256) k_hist_merge(const unsigned long long* __restrict__ chunk_hist, unsigned long long* hist, uint32_t n,
                                                    uint32_t copies, uint32_t* cursors) {
    if (cursors[3]) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = 0;
    for (uint32_t c = 0; c < copies; ++c) v += chunk_hist[(size_t)c * n + i];
    if (v) {
        atomicAdd(hist + i, v);
        if ((i & 3u) == 0u) atomicAdd(cursors + 4, (uint32_t)min(v, 0xFFFFFFFFull));  // labelled (non-U) rows of the chunk
    }
}
// single-pass ingest: labels[] in GAF row order from the record table.  tile_off = exclusive scan of tile_info[].z.
__global__ void __launch_bounds__(256) k_tile_rows(const uint4* __restrict__ tile_info, uint32_t* __restrict__ rows, uint32_t n_tiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_tiles) rows[t] = tile_info[t].z;
}
__global__ void __launch_bounds__(256) k_labels_from_table(const uint4* __restrict__ tile_info, const uint64_t* __restrict__ tile_off,
                                                           const uint4* __restrict__ meta_b, const uint16_t* __restrict__ row_key,
                                                           uint32_t* __restrict__ labels) {
    const uint4 ti = tile_info[blockIdx.x];
    const uint64_t base = tile_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < ti.y; i += blockDim.x) {
        const uint4 mb = meta_b[ti.x + i];
        if (mb.y & RM_VALID) labels[base + row_key[ti.x + i]] = mb.z;
    }
}

// =====================================================================================
// K7: unique trio table (profile.rs:658-740)
// =====================================================================================
__device__ __forceinline__ int64_t path_of_step(const uint64_t* __restrict__ poff, int64_t Htot, uint64_t k) {
    int64_t a = 0, b = Htot;  // last h with poff[h] <= k
    while (a < b) {
        int64_t m = (a + b) >> 1;
        if (poff[m] <= k) a = m + 1; else b = m;
    }
    return a - 1;
}

// Distinct-node marks of every path in ONE launch: a bit per (path, node of the path's species) - set with atomicOr; the visit
// that finds its bit already set is a repeat and gets bit 31 of its pnode entry (which of two visits of a node stays unmarked
// does not matter: the sums over distinct nodes are order-free).  pbm_off[h] = first bitmap word of path h, pbase[h] = first
// global node index of its species.  Round 1 launched one kernel per path rank (thousands at 5,000 paths per species).
__global__ void __launch_bounds__(256) k_mark_path_dups(uint32_t* pnode, const uint64_t* __restrict__ poff, int64_t Htot, int64_t P,
                                                        const uint64_t* __restrict__ pbm_off, const uint32_t* __restrict__ pbase, uint32_t* bm) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (uint64_t)P; k += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t h = path_of_step(poff, Htot, k);
        const uint32_t g = pnode[k] & 0x7FFFFFFFu;
        const uint32_t loc = g - pbase[h];
        const uint32_t bit = 1u << (loc & 31u);
        const uint32_t old = atomicOr(bm + pbm_off[h] + (loc >> 5), bit);
        if (old & bit) pnode[k] = g | 0x80000000u;  // node already met in this path
    }
}

__global__ void __launch_bounds__(256) k_path_len_sum(const uint32_t* __restrict__ pnode, const uint64_t* __restrict__ poff, int64_t Htot,
                                                      int64_t P, const uint32_t* __restrict__ val, unsigned long long* out) {
    // A block covers 256*16 consecutive steps (coalesced, strided by 256).  Steps of the block's first path
    // are reduced in shared memory; the (rare) steps of later paths go straight to their accumulators.
    constexpr int ITEMS = 16;
    const uint64_t base = (uint64_t)blockIdx.x * 256ull * ITEMS;
    __shared__ int64_t h0_s;
    __shared__ uint64_t h0_end_s;
    __shared__ unsigned long long ws[8];
    if (threadIdx.x == 0) {
        const int64_t h0 = path_of_step(poff, Htot, base);
        h0_s = h0;
        h0_end_s = poff[h0 + 1];
    }
    __syncthreads();
    const int64_t h0 = h0_s;
    const uint64_t h0_end = h0_end_s;
    unsigned long long acc = 0;
    // all 16 steps of a thread are loaded before any of their values is gathered, and all values before they are summed: 16
    // independent loads in flight per thread instead of load -> gather -> add chains (the kernel is latency-bound: 40 M steps)
    uint32_t g[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint64_t k = base + (uint64_t)i * 256ull + threadIdx.x;
        g[i] = k < (uint64_t)P ? __ldg(pnode + k) : 0x80000000u;  // bit 31: node already counted for this path (or no step)
    }
    uint32_t v[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) v[i] = (g[i] & 0x80000000u) ? 0u : __ldg(val + g[i]);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (g[i] & 0x80000000u) continue;
        const uint64_t k = base + (uint64_t)i * 256ull + threadIdx.x;
        if (k < h0_end) acc += v[i];
        else atomicAdd(out + path_of_step(poff, Htot, k), (unsigned long long)v[i]);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        if (t) atomicAdd(out + h0, t);
    }
}

__device__ __forceinline__ bool window_at(const uint32_t* __restrict__ pnode, const uint64_t* __restrict__ poff, int64_t Htot, uint64_t k,
                                          int64_t& h, uint32_t& lo, uint32_t& mid, uint32_t& hi) {
    h = path_of_step(poff, Htot, k);
    if (k + 2 >= poff[h + 1]) return false;
    const uint32_t x = pnode[k] & 0x7FFFFFFFu, y = pnode[k + 1] & 0x7FFFFFFFu, z = pnode[k + 2] & 0x7FFFFFFFu;
    lo = x < z ? x : z;
    hi = x < z ? z : x;
    mid = y;
    return true;
}

__device__ __forceinline__ uint4 ld_key(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// claim-or-find the slot of key (lo,mid,hi); empty slot = {TT_EMPTY,...}
__device__ __forceinline__ uint32_t key_slot(uint4* keys, uint32_t mask, uint32_t lo, uint32_t mid, uint32_t hi, bool insert) {
    uint32_t i = trio_hash(lo, mid, hi) & mask;
    const ulonglong2 empty = make_ulonglong2(0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull);
    const ulonglong2 mine = make_ulonglong2(((uint64_t)mid << 32) | lo, (uint64_t)hi);
    for (;;) {
        uint4 e = ld_key(keys + i);
        if (e.x == TT_EMPTY) {
            if (!insert) return 0xFFFFFFFFu;
            ulonglong2 prev = atomic_cas128(reinterpret_cast<ulonglong2*>(keys + i), empty, mine);
            if (prev.x == empty.x && prev.y == empty.y) return i;
            e.x = (uint32_t)prev.x;
            e.y = (uint32_t)(prev.x >> 32);
            e.z = (uint32_t)prev.y;
        }
        if (e.x == lo && e.y == mid && e.z == hi) return i;
        i = (i + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) k_trio_count(const uint32_t* __restrict__ pnode, const uint64_t* __restrict__ poff, int64_t Htot,
                                                    int64_t P, uint4* keys, uint32_t* cnt, uint32_t mask) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (uint64_t)P; k += (uint64_t)gridDim.x * blockDim.x) {
        int64_t h;
        uint32_t lo, mid, hi;
        if (!window_at(pnode, poff, Htot, k, h, lo, mid, hi)) continue;
        const uint32_t s = key_slot(keys, mask, lo, mid, hi, true);
        atomicAdd(cnt + s, 1u);  // occurrences with multiplicity over all paths (profile.rs:689-702)
    }
}

__global__ void __launch_bounds__(256) k_trio_flag(const uint32_t* __restrict__ pnode, const uint64_t* __restrict__ poff, int64_t Htot,
                                                   int64_t P, uint4* keys, const uint32_t* __restrict__ cnt, uint32_t mask,
                                                   uint32_t* __restrict__ flag) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (uint64_t)P; k += (uint64_t)gridDim.x * blockDim.x) {
        int64_t h;
        uint32_t lo, mid, hi;
        uint32_t f = 0;
        if (window_at(pnode, poff, Htot, k, h, lo, mid, hi)) {
            const uint32_t s = key_slot(keys, mask, lo, mid, hi, false);
            f = (s != 0xFFFFFFFFu && cnt[s] == 1u) ? 1u : 0u;  // profile.rs:709
        }
        flag[k] = f;
    }
}

// ---- exclusive scan uint32 -> uint64 (three kernels; block = 1024 threads x 2 items)
constexpr int SCAN_ITEMS = 2048;
__global__ void __launch_bounds__(1024) k_scan_block_sums(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ bsum) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_ITEMS;
    uint32_t v = 0;
    for (int i = 0; i < 2; ++i) {
        uint64_t k = base + threadIdx.x + i * 1024;
        if (k < n) v += in[k];
    }
    v = __reduce_add_sync(0xffffffffu, v);
    __shared__ uint32_t ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = __reduce_add_sync(0xffffffffu, ws[threadIdx.x]);
        if (threadIdx.x == 0) bsum[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(1024) k_scan_bsums(uint64_t* bsum, uint64_t nb) {
    // single block: exclusive scan of nb uint64 in place, bsum[nb] = total
    __shared__ uint64_t ws[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nb; base += 1024) {
        uint64_t i = base + threadIdx.x;
        uint64_t v = i < nb ? bsum[i] : 0;
        uint64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t s = ws[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t y = __shfl_up_sync(0xffffffffu, s, d);
                if (threadIdx.x >= d) s += y;
            }
            ws[threadIdx.x] = s;
        }
        __syncthreads();
        uint64_t carry = carry_s;
        uint64_t wbase = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
        if (i < nb) bsum[i] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry_s;
}
__global__ void __launch_bounds__(1024) k_scan_final(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ bsum,
                                                     uint64_t nb, uint64_t* __restrict__ out) {
    // block-local exclusive scan of 2048 items (thread t owns items 2t, 2t+1) + block offset
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_ITEMS;
    const uint64_t k0 = base + 2ull * threadIdx.x;
    uint32_t a = k0 < n ? in[k0] : 0, b = (k0 + 1) < n ? in[k0 + 1] : 0;
    uint32_t v = a + b, x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) >= d) x += y;
    }
    __shared__ uint32_t ws[32];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = ws[threadIdx.x];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (threadIdx.x >= d) s += y;
        }
        ws[threadIdx.x] = s;
    }
    __syncthreads();
    const uint32_t wbase = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const uint64_t ex = bsum[blockIdx.x] + wbase + x - v;
    if (k0 < n) out[k0] = ex;
    if (k0 + 1 < n) out[k0 + 1] = ex + a;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

__global__ void __launch_bounds__(256) k_trio_emit(const uint32_t* __restrict__ pnode, const uint64_t* __restrict__ poff, int64_t Htot,
                                                   int64_t P, const uint32_t* __restrict__ flag, const uint64_t* __restrict__ scan,
                                                   const uint32_t* __restrict__ len, uint32_t* __restrict__ trio_key,
                                                   int64_t* __restrict__ trio_len, uint32_t* __restrict__ trio_owner, uint4* tt,
                                                   uint32_t tt_mask, uint4* ninfo, uint64_t* __restrict__ trio_start) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < (uint64_t)P; k += (uint64_t)gridDim.x * blockDim.x) {
        if (!flag[k]) continue;
        int64_t h;
        uint32_t lo, mid, hi;
        window_at(pnode, poff, Htot, k, h, lo, mid, hi);
        const uint64_t t = scan[k];
        trio_key[3 * t + 0] = lo;
        trio_key[3 * t + 1] = mid;
        trio_key[3 * t + 2] = hi;
        trio_len[t] = (int64_t)len[lo] + (int64_t)len[mid] + (int64_t)len[hi];  // profile.rs:712
        trio_owner[t] = (uint32_t)h;
        atomicOr(&ninfo[mid].y, NI_TRIO_MID | trio_sig_bit(lo, hi));  // (lo < hi: window_at returns the canonical key)
        // unique keys: plain claim of an empty slot, then publish idx
        uint32_t i = trio_hash(lo, mid, hi) & tt_mask;
        const ulonglong2 empty = make_ulonglong2(0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull);
        const ulonglong2 mine = make_ulonglong2(((uint64_t)mid << 32) | lo, ((uint64_t)(uint32_t)t << 32) | hi);
        for (;;) {
            ulonglong2 prev = atomic_cas128(reinterpret_cast<ulonglong2*>(tt + i), empty, mine);
            if (prev.x == empty.x && prev.y == empty.y) break;
            i = (i + 1) & tt_mask;
        }
    }
    // first trio index of every hap (+ sentinel)
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h <= (uint64_t)Htot; h += (uint64_t)gridDim.x * blockDim.x)
        trio_start[h] = scan[poff[h]];
}

// =====================================================================================
// finalize kernels
// =====================================================================================
// profile.rs:844/:874 -> :1018-1023: covered bases per node = popcount of its bit range, or len if fully covered
__global__ void __launch_bounds__(256) k_cov(const uint32_t* __restrict__ len, const uint64_t* __restrict__ bit_off, const uint8_t* __restrict__ full,
                                             const uint32_t* __restrict__ bits, uint32_t* __restrict__ cov, int64_t N) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const uint32_t ln = len[g];
    if (full && full[g]) {
        cov[g] = ln;
        return;
    }
    const uint64_t b0 = bit_off[g], b1 = b0 + ln - 1;
    const uint64_t w0 = b0 >> 5, w1 = b1 >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u), m1 = 0xFFFFFFFFu >> (31u - (uint32_t)(b1 & 31u));
    uint32_t c;
    if (w0 == w1) {
        c = __popc(bits[w0] & m0 & m1);
    } else {
        c = __popc(bits[w0] & m0) + __popc(bits[w1] & m1);
        for (uint64_t w = w0 + 1; w < w1; ++w) c += __popc(bits[w]);
    }
    cov[g] = c;
}

__global__ void __launch_bounds__(256) k_hap_nz(const unsigned long long* __restrict__ trio_bases, const uint32_t* __restrict__ owner, int64_t T,
                                                unsigned long long* hap_nz) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool nz = t < T && (long long)trio_bases[t] > 0;  // abundance > 0 (profile.rs:1132)
    const uint32_t h = t < T ? owner[t] : 0xFFFFFFFFu;
    // trios of one hap are contiguous: aggregate lanes sharing the hap
    const unsigned peers = __match_any_sync(0xffffffffu, h);
    const unsigned votes = __ballot_sync(0xffffffffu, nz) & peers;
    if (t < T && votes && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hap_nz + h, (unsigned long long)__popc(votes));
}

__global__ void __launch_bounds__(256) k_depth(const unsigned long long* __restrict__ num, const uint32_t* __restrict__ den32,
                                               const int64_t* __restrict__ den64, double* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = den32 ? (double)den32[i] : (double)den64[i];
    out[i] = (double)(long long)num[i] / d;  // profile.rs:988 / :1014 (IEEE division, identical on host)
}

// multi-GPU finalize: a node some read covered completely carries NI_FULL instead of bits.  Before the bitmaps of the ranks are
// OR-ed its bits are written out, so that the bitmap alone carries the coverage and the flags need no reduction of their own.
__global__ void __launch_bounds__(256) k_bits_fill_full(const uint4* __restrict__ ninfo, uint32_t* bits, int64_t N) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    const uint4 ni = ninfo[g];
    if (!(ni.y & NI_FULL)) return;
    const uint64_t b0 = ((uint64_t)ni.w << 32) | ni.z, b1 = b0 + ni.x - 1;
    const uint64_t w0 = b0 >> 5, w1 = b1 >> 5;
    const uint32_t m0 = 0xFFFFFFFFu << (b0 & 31u), m1 = 0xFFFFFFFFu >> (31u - (uint32_t)(b1 & 31u));
    if (w0 == w1) { atomicOr(bits + w0, m0 & m1); return; }
    atomicOr(bits + w0, m0);  // the first and the last word are shared with the neighbours
    for (uint64_t w = w0 + 1; w < w1; ++w) bits[w] = 0xFFFFFFFFu;
    atomicOr(bits + w1, m1);
}
// dst[i] |= src[0][i] | src[1][i] | ... (n_src slices of n words behind each other)
__global__ void __launch_bounds__(256) k_or_slices(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n_src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v = dst[i];
        for (uint32_t q = 0; q < n_src; ++q) v |= src[(uint64_t)q * n + i];
        dst[i] = v;
    }
}
__global__ void __launch_bounds__(256) k_or_words(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] |= src[i];
}

__global__ void __launch_bounds__(256) k_ninfo_build(const uint32_t* __restrict__ len, const uint64_t* __restrict__ bit_off, uint4* __restrict__ ninfo, int64_t N) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < N) ninfo[g] = make_uint4(len[g], 0u, (uint32_t)bit_off[g], (uint32_t)(bit_off[g] >> 32));
}
// mode 0: full[g] = NI_FULL flag of ninfo (finalize / cross-rank max-reduce staging); mode 1: clear the flag (reset)
__global__ void __launch_bounds__(256) k_ninfo_full(uint4* ninfo, uint8_t* __restrict__ full, int64_t N, int mode) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    if (mode == 0) full[g] = (uint8_t)(ninfo[g].y & NI_FULL);
    else ninfo[g].y &= ~NI_FULL;
}


// =====================================================================================
// K10: long-read best-alignment filter (gaf_filter.rs:22-97)
// =====================================================================================
// newlines per 4 KB micro-tile (one warp each)
__global__ void __launch_bounds__(256) k_flt_count_nl(const uint8_t* __restrict__ text, uint64_t n, uint32_t n_micro, uint32_t* __restrict__ cnt) {
    const uint32_t mt = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (mt >= n_micro) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t base = (uint64_t)mt * MICRO;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < (int)(MICRO / 512); ++j) {
        const uint64_t off = base + (uint64_t)j * 512u + lane * 16u;
        if (off >= n) break;
        uint4 q = __ldg(reinterpret_cast<const uint4*>(text + off));
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t m = nl_mask4(w[i]);
            const uint64_t wo = off + i * 4;
            if (wo + 4 > n) {  // the last word may reach into the padding
                uint32_t keep = 0;
                for (int k = 0; k < 4; ++k)
                    if (wo + k < n) keep |= 0x80u << (8 * k);
                m &= keep;
            }
            c += __popc(m);
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) cnt[mt] = c;
}
// line i+1 starts behind the i-th newline; line 0 starts at 0
__global__ void __launch_bounds__(256) k_flt_line_starts(const uint8_t* __restrict__ text, uint64_t n, uint32_t n_micro,
                                                         const uint64_t* __restrict__ micro_base, uint64_t* __restrict__ line_off) {
    const uint32_t mt = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (mt >= n_micro) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t base = (uint64_t)mt * MICRO;
    uint64_t idx = micro_base[mt];
    if (mt == 0 && lane == 0) line_off[0] = 0;
    for (int j = 0; j < (int)(MICRO / 512); ++j) {
        const uint64_t off = base + (uint64_t)j * 512u + lane * 16u;
        uint32_t c = 0;
        uint32_t msk[4] = {0, 0, 0, 0};
        if (off < n) {
            uint4 q = __ldg(reinterpret_cast<const uint4*>(text + off));
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t m = nl_mask4(w[i]);
                const uint64_t wo = off + i * 4;
                if (wo + 4 > n) {
                    uint32_t keep = 0;
                    for (int k = 0; k < 4; ++k)
                        if (wo + k < n) keep |= 0x80u << (8 * k);
                    m &= keep;
                }
                msk[i] = m;
                c += __popc(m);
            }
        }
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        uint64_t my = idx + x - c;
        idx += __shfl_sync(0xffffffffu, x, 31);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t m = msk[i];
            while (m) {
                uint32_t byte = (__ffs(m) - 1) >> 3;
                m &= m - 1;
                line_off[++my] = off + i * 4 + byte + 1;
            }
        }
    }
}

struct FltRec {  // one parsed line (gaf_filter.rs:12-20)
    ulonglong2 key;   // read-id hash, {0,0} = line rejected by parse_line
    long long matches;
    double ident;
    int mapq, span;
};

__device__ __forceinline__ bool flt_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// Rust `str::parse::<i32>`: [+-]?digits, no whitespace, must fit i32
__device__ bool flt_parse_i32(const uint8_t* p, uint32_t n, int& out) {
    if (n == 0) return false;
    uint32_t i = 0;
    bool neg = false;
    if (p[0] == '+' || p[0] == '-') { neg = p[0] == '-'; i = 1; }
    if (i == n) return false;
    long long v = 0;
    for (; i < n; ++i) {
        uint32_t d = (uint32_t)p[i] - '0';
        if (d > 9u) return false;
        v = v * 10 + d;
        if (v > 2147483648ll) return false;
    }
    v = neg ? -v : v;
    if (v < -2147483648ll || v > 2147483647ll) return false;
    out = (int)v;
    return true;
}

// Decimal -> double for [+-]?digits[.digits][e[+-]digits] with <= 15 significant digits and a decimal exponent in
// [-22, 22]: both factors are exact doubles, so one IEEE multiply/divide is correctly rounded (the same value
// Rust's f64::from_str returns).  status: 0 ok, 1 not a number (line rejected), 2 outside the exact range.
__device__ int flt_parse_f64(const uint8_t* p, uint32_t n, double& out) {
    const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    if (n == 0) return 1;
    uint32_t i = 0;
    bool neg = false;
    if (p[0] == '+' || p[0] == '-') { neg = p[0] == '-'; i = 1; }
    unsigned long long mant = 0;
    int sig = 0, exp10 = 0, nd = 0;
    bool dot = false;
    for (; i < n; ++i) {
        uint8_t c = p[i];
        uint32_t d = (uint32_t)c - '0';
        if (d <= 9u) {
            ++nd;
            if (mant == 0 && d == 0) { if (dot) --exp10; continue; }  // leading zeros
            if (sig < 19) { mant = mant * 10 + d; ++sig; if (dot) --exp10; }
            else { if (d != 0) return 2; if (!dot) ++exp10; }
        } else if (c == '.' && !dot) {
            dot = true;
        } else {
            break;
        }
    }
    if (nd == 0) {  // Rust also accepts inf / infinity / nan (any case): representable, but outside this exact path
        const uint32_t rem = n - i;
        auto lc = [&](uint32_t k) { return (uint8_t)(p[i + k] | 0x20); };
        if (!dot && rem == 3 && ((lc(0) == 'i' && lc(1) == 'n' && lc(2) == 'f') || (lc(0) == 'n' && lc(1) == 'a' && lc(2) == 'n'))) return 2;
        if (!dot && rem == 8 && lc(0) == 'i' && lc(1) == 'n' && lc(2) == 'f' && lc(3) == 'i' && lc(4) == 'n' && lc(5) == 'i' && lc(6) == 't' && lc(7) == 'y') return 2;
        return 1;
    }
    if (i < n) {
        if (p[i] != 'e' && p[i] != 'E') return 1;
        ++i;
        bool eneg = false;
        if (i < n && (p[i] == '+' || p[i] == '-')) { eneg = p[i] == '-'; ++i; }
        if (i == n) return 1;
        int e = 0;
        for (; i < n; ++i) {
            uint32_t d = (uint32_t)p[i] - '0';
            if (d > 9u) return 1;
            if (e < 100000) e = e * 10 + (int)d;
        }
        exp10 += eneg ? -e : e;
    }
    if (mant == 0) { out = neg ? -0.0 : 0.0; return 0; }
    while (sig > 15 && mant % 10 == 0) { mant /= 10; ++exp10; --sig; }
    if (sig > 15 || exp10 < -22 || exp10 > 22) return 2;
    double v = (double)mant;
    v = exp10 >= 0 ? v * P10[exp10] : v / P10[-exp10];
    out = neg ? -v : v;
    return 0;
}

__global__ void __launch_bounds__(128) k_flt_parse(const uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ line_off,
                                                   uint64_t n_lines, FltRec* __restrict__ recs, uint32_t* flags) {
    const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    FltRec r;
    r.key = make_ulonglong2(0ull, 0ull);
    r.matches = 0; r.ident = 0; r.mapq = 0; r.span = 0;
    uint64_t a = line_off[li], b = line_off[li + 1] - 1;  // [a,b): the line without its '\n'
    while (a < b && flt_ws(text[a])) ++a;                 // gaf_filter.rs:23 line.trim()
    while (b > a && flt_ws(text[b - 1])) --b;
    // fields 1..16: [fs[k], fe[k])
    uint64_t fs[16], fe[16];
    int nf = 0;
    uint64_t s0 = a;
    for (uint64_t k = a; k <= b; ++k) {
        if (k == b || text[k] == '\t') {
            if (nf < 16) { fs[nf] = s0; fe[nf] = k; }
            ++nf;
            s0 = k + 1;
        }
    }
    if (nf >= 16) {
        int m, q, c4, c3;
        bool ok = flt_parse_i32(text + fs[9], (uint32_t)(fe[9] - fs[9]), m);
        uint64_t c = fe[15];
        while (c > fs[15] && text[c - 1] != ':') --c;  // rsplit(':').next(): text behind the last ':'
        double id = 0;
        int fst = ok ? flt_parse_f64(text + c, (uint32_t)(fe[15] - c), id) : 1;
        if (fst == 2) atomicOr(flags + 2, 1u);
        ok = ok && fst == 0 && flt_parse_i32(text + fs[11], (uint32_t)(fe[11] - fs[11]), q) &&
             flt_parse_i32(text + fs[3], (uint32_t)(fe[3] - fs[3]), c4) && flt_parse_i32(text + fs[2], (uint32_t)(fe[2] - fs[2]), c3);
        if (ok) {
            IdHasher H;
            for (uint64_t k = fs[0]; k < fe[0]; ++k) H.byte(text[k]);
            IdHash h = H.finish();
            r.key = make_ulonglong2(h.lo, ((uint64_t)h.hi << 32) | 1u);
            r.matches = m;
            r.ident = id;
            r.mapq = q;
            r.span = c4 - c3;  // i32 wrap-around is a debug-only panic in Rust; release builds wrap
        }
    }
    recs[li] = r;
}

__device__ __forceinline__ uint64_t flt_slot(ulonglong2* keys, uint64_t mask, uint32_t shift, const ulonglong2& key) {
    IdHash h;
    h.lo = key.x;
    h.hi = (uint32_t)(key.y >> 32);
    uint64_t i = ds_home(h, shift);
    for (;;) {
        ulonglong2 cur = ld128(keys + i);
        if (cur.x == 0ull && cur.y == 0ull) {
            cur = atomic_cas128(keys + i, make_ulonglong2(0ull, 0ull), key);
            if (cur.x == 0ull && cur.y == 0ull) return i;
        }
        if (cur.x == key.x && cur.y == key.y) return i;
        i = (i + 1) & mask;
    }
}
// best (matches, identity) per read id over ALL parsed lines (gaf_filter.rs:65-74)
__global__ void __launch_bounds__(256) k_flt_best(const FltRec* __restrict__ recs, uint64_t n_lines, ulonglong2* keys, ulonglong2* best,
                                                  uint64_t mask, uint32_t shift) {
    const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    const FltRec r = recs[li];
    if (r.key.x == 0ull && r.key.y == 0ull) return;
    const uint64_t i = flt_slot(keys, mask, shift, r.key);
    // slot value {matches + 2^40 (0 = unset), identity bits}
    const ulonglong2 mine = make_ulonglong2((unsigned long long)(r.matches + (1ll << 40)), (unsigned long long)__double_as_longlong(r.ident));
    ulonglong2 cur = ld128(best + i);
    for (;;) {
        bool better = cur.x == 0ull;
        if (!better) {
            const long long cm = (long long)cur.x - (1ll << 40);
            const double ci = __longlong_as_double((long long)cur.y);
            better = r.matches > cm || (r.matches == cm && r.ident > ci);
        }
        if (!better) break;
        const ulonglong2 prev = atomic_cas128(best + i, cur, mine);
        if (prev.x == cur.x && prev.y == cur.y) break;
        cur = prev;
    }
}
// first qualifying line per id in file order (the reference's choice is a race, gaf_filter.rs:80-93)
__global__ void __launch_bounds__(256) k_flt_first(const FltRec* __restrict__ recs, uint64_t n_lines, ulonglong2* keys,
                                                   const ulonglong2* __restrict__ best, unsigned long long* first, uint64_t mask, uint32_t shift,
                                                   uint8_t* __restrict__ qual) {
    const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    const FltRec r = recs[li];
    bool q = false;
    if (!(r.key.x == 0ull && r.key.y == 0ull) && r.mapq > 20 && r.span > 1000) {
        const uint64_t i = flt_slot(keys, mask, shift, r.key);
        const ulonglong2 b = best[i];
        q = ((long long)b.x - (1ll << 40)) == r.matches && __longlong_as_double((long long)b.y) == r.ident;
        if (q) atomicMin(first + i, (unsigned long long)li);
    }
    qual[li] = q ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_flt_select(const FltRec* __restrict__ recs, uint64_t n_lines, ulonglong2* keys,
                                                    const unsigned long long* __restrict__ first, uint64_t mask, uint32_t shift,
                                                    const uint8_t* __restrict__ qual, uint32_t* __restrict__ sel) {
    const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines) return;
    uint32_t s = 0;
    if (qual[li]) {
        const uint64_t i = flt_slot(keys, mask, shift, recs[li].key);
        s = first[i] == li ? 1u : 0u;
    }
    sel[li] = s;
}
__global__ void __launch_bounds__(256) k_flt_compact(const uint32_t* __restrict__ sel, const uint64_t* __restrict__ scan,
                                                     const uint64_t* __restrict__ line_off, uint64_t n_lines, uint64_t* __restrict__ out, uint64_t cap) {
    const uint64_t li = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lines || !sel[li]) return;
    const uint64_t j = scan[li];
    if (j < cap) out[j] = line_off[li];
}

// =====================================================================================
// GFA graph text -> node lengths + path steps on the device (profile.rs:466-545 read_gfa, previous = 0): SURVEY section 8(f3)
//   k_gfa_lines        one thread per line: 'S' lines write len[id - 1] (the reference asserts that the S lines come in id order
//                      without gaps: k_gfa_check_order verifies id - 1 == number of S lines before it, from a scan of the S flags);
//                      'P' / 'W' lines are appended to a list {line, haplotype-name extent, path-field extent}
//   k_gfa_path_count   the path fields are cut into 4 KB pieces (host table): ids ending in each piece (a digit followed by a non-digit)
//   k_gfa_path_decode  ... and written, as id - 1, at the piece's destination + rank (the host lays the lines of one haplotype out
//                      behind each other in file order - "chromosomes of one genome merge into one path", profile.rs:536-541)
// Error bits (flags[0]): 1 S id not a number / out of range, 2 S lines out of order, 4 node length 0, 8 path line without a name field,
// 16 path node id outside the graph, 32 a tab inside a W line's walk field (more than seven fields).
// =====================================================================================
struct GfaPathLine { unsigned long long line, name_beg, fld_beg, fld_end; uint32_t name_len, kind, pad0, pad1; };  // 48 bytes
constexpr uint32_t GFA_PIECE = 4096;

__device__ __forceinline__ bool gfa_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }  // what str::trim removes (ASCII)

__global__ void __launch_bounds__(256) k_gfa_lines(const uint8_t* __restrict__ text, const uint64_t* __restrict__ line_off, uint64_t n_lines,
                                                   int64_t n_nodes, uint32_t* __restrict__ len, uint32_t* __restrict__ is_s,
                                                   uint32_t* __restrict__ s_adj, GfaPathLine* __restrict__ plist, unsigned long long* pcount,
                                                   uint64_t pcap, uint32_t* flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const uint64_t b = line_off[i];
    uint64_t e = line_off[i + 1] - 1;  // the '\n'
    is_s[i] = 0;
    if (e <= b) return;
    const uint8_t c0 = text[b];
    if (c0 != 'S' && c0 != 'P' && c0 != 'W') return;
    while (e > b && gfa_space(text[e - 1])) --e;  // line.trim() (the line starts with a letter: nothing to trim in front)
    // the first tabs (three; six for a W line, whose walk is its 7th and last field)
    uint64_t t[6] = {e, e, e, e, e, e};
    int nt = 0;
    const int want = c0 == 'W' ? 6 : (c0 == 'P' ? 2 : 3);  // (the path field of a P line is megabytes long: its end is found by k_gfa_field_end)
    for (uint64_t p = b; p < e && nt < want; ++p)
        if (text[p] == '\t') t[nt++] = p;
    if (c0 == 'S') {
        if (nt < 2) return;  // parts.len() < 3: skipped
        // parts[1].parse::<usize>(): digits only (a leading '+' is accepted by Rust's parser)
        uint64_t p = t[0] + 1, v = 0;
        uint32_t nd = 0;
        if (p < t[1] && text[p] == '+') ++p;
        for (; p < t[1]; ++p) {
            const uint32_t d = (uint32_t)text[p] - (uint32_t)'0';
            if (d > 9u || nd >= 18u) { nd = 0xFFFFu; break; }
            v = v * 10u + d;
            ++nd;
        }
        if (nd == 0u || nd == 0xFFFFu || v == 0 || (int64_t)(v - 1) >= n_nodes) { atomicOr(flags, 1u); return; }
        const uint64_t seq_end = nt >= 3 ? t[2] : e;
        const uint64_t l = seq_end - (t[1] + 1);
        if (l == 0) { atomicOr(flags, 4u); return; }
        len[v - 1] = (uint32_t)(l > 0xFFFFFFFFull ? 0xFFFFFFFFull : l);
        is_s[i] = 1;
        s_adj[i] = (uint32_t)(v - 1);
        return;
    }
    // 'P' / 'W'
    if (nt < 1) { atomicOr(flags, 8u); return; }  // parts[1] does not exist: the reference panics
    const bool is_w = (t[0] == b + 1) && c0 == 'W';  // parts[0] == "W"
    GfaPathLine pl;
    pl.line = i;
    pl.kind = is_w ? 2u : 1u;
    pl.pad0 = pl.pad1 = 0;
    pl.name_beg = t[0] + 1;
    uint64_t name_end = t[1];  // (== e when the line has two fields)
    if (!is_w)
        for (uint64_t p = pl.name_beg; p < name_end; ++p)
            if (text[p] == '#') { name_end = p; break; }  // parts[1].split('#').next()
    pl.name_len = (uint32_t)(name_end - pl.name_beg);
    if (is_w) {
        // parts.last(): a W line is `W sample hap seq start end walk` - the walk (megabytes long) is what follows the sixth tab; that
        // no tab hides further inside it (a line with more fields) is checked by the piece kernels, not by a serial scan here.  A W
        // line with fewer fields: the last field starts behind its last tab.
        pl.fld_beg = nt ? t[nt - 1] + 1 : b;
        pl.fld_end = e;
    } else {     // parts.get(2) or "": from behind the second tab to the next tab (k_gfa_field_end) or the end of the line
        pl.fld_beg = nt >= 2 ? t[1] + 1 : e;
        pl.fld_end = e;
    }
    const unsigned long long k = atomicAdd(pcount, 1ull);
    if (k < pcap) plist[k] = pl;
}
__global__ void __launch_bounds__(256) k_gfa_check_order(const uint32_t* __restrict__ is_s, const uint32_t* __restrict__ s_adj,
                                                         const uint64_t* __restrict__ ord, uint64_t n_lines, uint32_t* flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lines && is_s[i] && (uint64_t)s_adj[i] != ord[i]) atomicOr(flags, 2u);  // "Node ID out of order or mismatch" (profile.rs:489)
}
// first tab inside the candidate path field of every P / W line (pieces of 4 KB, as below): the P field ends there; a W line has more
// than seven fields then
__global__ void __launch_bounds__(256) k_gfa_field_end(const uint8_t* __restrict__ text, const uint64_t* __restrict__ piece_beg,
                                                       const uint64_t* __restrict__ piece_fend, const uint32_t* __restrict__ piece_line,
                                                       unsigned long long* line_tab) {
    const uint64_t pb = piece_beg[blockIdx.x], fe = piece_fend[blockIdx.x];
    const uint64_t pe = pb + GFA_PIECE < fe ? pb + GFA_PIECE : fe;
    const uint64_t p0 = pb + 16ull * threadIdx.x;
    for (uint32_t k = 0; k < 16u && p0 + k < pe; ++k)
        if (text[p0 + k] == '\t') { atomicMin(line_tab + piece_line[blockIdx.x], (unsigned long long)(p0 + k)); break; }
}
// ids ending in [pb, pe) of a path field that ends at fe: a digit whose successor (inside the field, or the byte that closes it) is none
__device__ __forceinline__ uint32_t gfa_ends16(const uint8_t* __restrict__ text, uint64_t p0, uint64_t pe, uint64_t fe) {
    uint32_t m = 0;
    bool d = p0 < pe && ((uint32_t)text[p0] - (uint32_t)'0') <= 9u;
    for (uint32_t k = 0; k < 16u && p0 + k < pe; ++k) {
        const bool dn = (p0 + k + 1 < fe) && ((uint32_t)text[p0 + k + 1] - (uint32_t)'0') <= 9u;
        if (d && !dn) m |= 1u << k;
        d = dn;
    }
    return m;
}
__global__ void __launch_bounds__(256) k_gfa_path_count(const uint8_t* __restrict__ text, const uint64_t* __restrict__ piece_beg,
                                                        const uint64_t* __restrict__ piece_fend, uint32_t* __restrict__ piece_cnt, uint32_t* flags) {
    __shared__ uint32_t wsum[8];
    const uint64_t pb = piece_beg[blockIdx.x], fe = piece_fend[blockIdx.x];
    const uint64_t pe = pb + GFA_PIECE < fe ? pb + GFA_PIECE : fe;
    const uint64_t p0 = pb + 16ull * threadIdx.x;
    uint32_t c = p0 < pe ? __popc(gfa_ends16(text, p0, pe, fe)) : 0u;
    (void)flags;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31u) == 0u) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        piece_cnt[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) k_gfa_path_decode(const uint8_t* __restrict__ text, const uint64_t* __restrict__ piece_beg,
                                                         const uint64_t* __restrict__ piece_fend, const uint64_t* __restrict__ piece_fbeg,
                                                         const uint64_t* __restrict__ piece_dst, int64_t n_nodes, uint32_t* __restrict__ out,
                                                         uint32_t* flags) {
    __shared__ uint32_t wsum[8];
    const uint64_t pb = piece_beg[blockIdx.x], fe = piece_fend[blockIdx.x], fb = piece_fbeg[blockIdx.x];
    const uint64_t pe = pb + GFA_PIECE < fe ? pb + GFA_PIECE : fe;
    const uint64_t p0 = pb + 16ull * threadIdx.x;
    uint32_t m = p0 < pe ? gfa_ends16(text, p0, pe, fe) : 0u;
    const uint32_t c = __popc(m), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31u) wsum[warp] = x;
    __syncthreads();
    uint32_t rank = x - c;
    for (uint32_t w = 0; w < warp; ++w) rank += wsum[w];
    uint64_t dst = piece_dst[blockIdx.x] + rank;
    while (m) {
        const uint32_t k = __ffs(m) - 1;
        m &= m - 1u;
        uint64_t q = p0 + k;  // last digit of the id; its first digit is at most 18 bytes in front, not before the field
        uint64_t v = 0, mul = 1;
        uint32_t nd = 0;
        while (q + 1 > fb && ((uint32_t)text[q] - (uint32_t)'0') <= 9u && nd < 19u) {
            v += (uint64_t)(text[q] - '0') * mul;
            mul *= 10u;
            ++nd;
            if (q == fb) break;
            --q;
        }
        if (v == 0 || (int64_t)(v - 1) >= n_nodes || nd >= 19u) { atomicOr(flags, 16u); v = 1; }
        out[dst++] = (uint32_t)(v - 1);
    }
}

// =====================================================================================
// launchers
// =====================================================================================
static inline uint32_t grid_for(uint64_t n, uint32_t per_block, uint32_t cap = 148u * 32u) {
    uint64_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (uint32_t)g;
}

void launch_gfa_lines(const uint8_t* text, const uint64_t* line_off, uint64_t n_lines, int64_t n_nodes, uint32_t* len, uint32_t* is_s, uint32_t* s_adj,
                      void* plist, unsigned long long* pcount, uint64_t pcap, uint32_t* flags, cudaStream_t st) {
    if (n_lines == 0) return;
    k_gfa_lines<<<(uint32_t)((n_lines + 255) / 256), 256, 0, st>>>(text, line_off, n_lines, n_nodes, len, is_s, s_adj, reinterpret_cast<GfaPathLine*>(plist),
                                                                      pcount, pcap, flags);
    PTX_LAUNCHED();
}
void launch_gfa_check_order(const uint32_t* is_s, const uint32_t* s_adj, const uint64_t* ord, uint64_t n_lines, uint32_t* flags, cudaStream_t st) {
    if (n_lines == 0) return;
    k_gfa_check_order<<<(uint32_t)((n_lines + 255) / 256), 256, 0, st>>>(is_s, s_adj, ord, n_lines, flags);
    PTX_LAUNCHED();
}
void launch_gfa_path_count(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, uint32_t* piece_cnt, uint32_t n_pieces, uint32_t* flags,
                           cudaStream_t st) {
    if (n_pieces == 0) return;
    k_gfa_path_count<<<n_pieces, 256, 0, st>>>(text, piece_beg, piece_fend, piece_cnt, flags);
    PTX_LAUNCHED();
}
void launch_gfa_path_decode(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, const uint64_t* piece_fbeg, const uint64_t* piece_dst,
                            int64_t n_nodes, uint32_t* out, uint32_t n_pieces, uint32_t* flags, cudaStream_t st) {
    if (n_pieces == 0) return;
    k_gfa_path_decode<<<n_pieces, 256, 0, st>>>(text, piece_beg, piece_fend, piece_fbeg, piece_dst, n_nodes, out, flags);
    PTX_LAUNCHED();
}
void launch_gfa_field_end(const uint8_t* text, const uint64_t* piece_beg, const uint64_t* piece_fend, const uint32_t* piece_line, unsigned long long* line_tab,
                          uint32_t n_pieces, cudaStream_t st) {
    if (n_pieces == 0) return;
    k_gfa_field_end<<<n_pieces, 256, 0, st>>>(text, piece_beg, piece_fend, piece_line, line_tab);
    PTX_LAUNCHED();
}
size_t gfa_path_line_bytes() { return sizeof(GfaPathLine); }
void launch_count_records(const uint8_t* text, uint64_t n_bytes, uint32_t n_micro, uint32_t* micro_count, unsigned long long* total_slots,
                          cudaStream_t st) {
    k_count_records<<<(n_micro + 7) / 8, 256, 0, st>>>(text, n_bytes, n_micro, micro_count, total_slots);
    PTX_LAUNCHED();
}

size_t ingest_smem_bytes(uint32_t tile_bytes, uint32_t over_bytes, bool short_kernel, bool multi_species) {
    const size_t hist_bytes = multi_species ? HIST_SLOTS * 5 * sizeof(uint32_t) : 0;
    if (short_kernel) {
        const size_t stage_bytes = (size_t)tile_bytes + over_bytes;
        return stage_bytes + STAGE_PAD + 2 * ((stage_bytes / 32 + 2 + 3) / 4 * 4) * sizeof(uint32_t) + SHORT_STASH_CAP * SHORT_THREADS * sizeof(uint32_t) +
               3 * SHORT_REC_CAP * sizeof(uint16_t);
    }
    return (size_t)tile_bytes + OVER + 16 + STASH_CAP * INGEST_THREADS * sizeof(uint32_t) + 4 * REC_CAP * sizeof(uint16_t) + hist_bytes;
}

size_t ingest_l_smem_bytes(uint32_t tile_bytes, uint32_t over_bytes, bool multi_species) {
    const size_t stage_bytes = (size_t)tile_bytes + over_bytes;
    const size_t bm_alloc = (stage_bytes / 32 + 2 + 3) / 4 * 4;
    return stage_bytes + STAGE_PAD + 4 * bm_alloc * sizeof(uint32_t) + bm_alloc * sizeof(uint16_t) + LONG_THREADS * (2 * sizeof(uint32_t) + 2 * sizeof(uint16_t)) +
           2 * LONG_REC_CAP * sizeof(uint16_t);
}

void launch_ingest(const IngestArgs& a, cudaStream_t st) {
    // the opt-in shared-memory size is a per-device attribute of the function
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaFuncSetAttribute(k_ingest<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ingest_smem_bytes(MAX_TILE, OVER, false, true));
        cudaFuncSetAttribute(k_ingest<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ingest_smem_bytes(MAX_TILE, OVER, false, true));
        cudaFuncSetAttribute(k_ingest_s<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ingest_smem_bytes(MAX_TILE, OVER, true, true) + 8192);
        cudaFuncSetAttribute(k_ingest_s<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ingest_smem_bytes(MAX_TILE, OVER, true, true) + 8192);
        cudaFuncSetAttribute(k_ingest_l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ingest_l_smem_bytes(MAX_TILE, OVER, true) + 8192);
        // all of the SM's shared memory for these kernels: the driver's own carve-out choice flips between launches of the
        // same configuration (measured: k_ingest_s 1.06 or 1.34 ms), and their occupancy is set by shared memory
        static const int carve = getenv("PTX_CARVEOUT") ? atoi(getenv("PTX_CARVEOUT")) : 100;
        if (carve >= 0) {
            cudaFuncSetAttribute(k_ingest<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_ingest<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_ingest_s<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_ingest_s<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_ingest_l, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        }
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const bool multi = a.ranges.S > 1;
    if (a.long_mode && !a.long_new) k_ingest<true><<<a.n_tiles, INGEST_THREADS, ingest_smem_bytes(a.tile_bytes, OVER, false, multi), st>>>(a);
    else if (a.long_mode) {
        static const int padl = getenv("PTX_SMEM_PAD_L") ? atoi(getenv("PTX_SMEM_PAD_L")) : 0;  // measurement knob: fewer resident CTAs
        k_ingest_l<<<a.n_tiles, LONG_THREADS, ingest_l_smem_bytes(a.tile_bytes, a.over_bytes, multi) + (size_t)std::min(padl, 8192), st>>>(a);
    }
    else if (a.old_short) k_ingest<false><<<a.n_tiles, INGEST_THREADS, ingest_smem_bytes(a.tile_bytes, OVER, false, multi), st>>>(a);
    else {
        static const int pad = getenv("PTX_SMEM_PAD") ? atoi(getenv("PTX_SMEM_PAD")) : 0, cls = getenv("PTX_CLS_MUL") ? atoi(getenv("PTX_CLS_MUL")) : 0;  // measurement knobs
        const size_t sm = ingest_smem_bytes(a.tile_bytes, a.over_bytes, true, multi) + (size_t)std::min(pad, 8192);
        if (cls) k_ingest_s<true><<<a.n_tiles, SHORT_THREADS, sm, st>>>(a);
        else k_ingest_s<false><<<a.n_tiles, SHORT_THREADS, sm, st>>>(a);
    }
    PTX_LAUNCHED();
}
// multi-GPU finalize: min(non-U rows of this rank, 1000) = sum of the species read counts, OR-ed into bits 8.. of the word whose
// bit 0 is the outbox-overflow marker (all-gathered with the box fills)
__global__ void __launch_bounds__(256) k_count_labelled(const unsigned long long* __restrict__ hist, uint32_t S, unsigned long long* dst) {
    __shared__ unsigned long long part[256];
    unsigned long long s = 0;
    for (uint32_t i = threadIdx.x; i < S; i += 256u) s += hist[4ull * i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t d = 128; d; d >>= 1) {
        if (threadIdx.x < d) part[threadIdx.x] += part[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // (ptx_finalize may run again after more input: replace the count, keep the marker bit)
        atomicAnd(dst, 0xFFull);
        atomicOr(dst, (part[0] < 1000ull ? part[0] : 1000ull) << 8);
    }
}
void launch_count_labelled(const unsigned long long* hist, uint32_t S, unsigned long long* dst, cudaStream_t st) {
    k_count_labelled<<<1, 256, 0, st>>>(hist, S, dst);
    PTX_LAUNCHED();
}
void launch_hist_merge(const unsigned long long* chunk_hist, unsigned long long* hist, uint32_t n, uint32_t copies, uint32_t* cursors, cudaStream_t st) {
    k_hist_merge<<<(n + 255u) / 256u, 256, 0, st>>>(chunk_hist, hist, n, copies, cursors);
    PTX_LAUNCHED();
}
void launch_tile_rows(const uint4* tile_info, uint32_t* rows, uint32_t n_tiles, cudaStream_t st) {
    k_tile_rows<<<(n_tiles + 255u) / 256u, 256, 0, st>>>(tile_info, rows, n_tiles);
    PTX_LAUNCHED();
}
void launch_labels_from_table(const uint4* tile_info, const uint64_t* tile_off, const uint4* meta_b, const uint16_t* row_key, uint32_t* labels,
                              uint32_t n_tiles, cudaStream_t st) {
    if (n_tiles == 0) return;
    k_labels_from_table<<<n_tiles, 256, 0, st>>>(tile_info, tile_off, meta_b, row_key, labels);
    PTX_LAUNCHED();
}
void launch_apply(const IngestArgs& a, uint32_t n_entries, int mode, cudaStream_t st) {
    if (n_entries == 0) return;
    const uint32_t grid = ((n_entries == ENTRIES_FROM_DEVICE ? a.slots_cap : n_entries) + 255u) / 256u;
    if (grid == 0) return;
#define PTX_APPLY_VARIANTS(M)                                                                   \
    switch (a.scatter_var) {                                                                   \
        case 1: k_apply<M, 1><<<grid, 256, 0, st>>>(a, n_entries); break;                       \
        case 2: k_apply<M, 2><<<grid, 256, 0, st>>>(a, n_entries); break;                       \
        case 3: k_apply<M, 3><<<grid, 256, 0, st>>>(a, n_entries); break;                       \
        default: k_apply<M, 0><<<grid, 256, 0, st>>>(a, n_entries); break;                      \
    }
    switch (mode) {
        case MODE_CLASSIFY: k_apply<MODE_CLASSIFY><<<grid, 256, 0, st>>>(a, n_entries); break;
        case MODE_CLASSIFY | MODE_COVER: PTX_APPLY_VARIANTS(MODE_CLASSIFY | MODE_COVER) break;
        case MODE_COVER: PTX_APPLY_VARIANTS(MODE_COVER) break;
        case MODE_COVER | MODE_KEEPMASK: PTX_APPLY_VARIANTS(MODE_COVER | MODE_KEEPMASK) break;
        case MODE_REBOX: k_apply<MODE_REBOX><<<grid, 256, 0, st>>>(a, n_entries); break;
        default: return;
    }
    PTX_LAUNCHED();
}
// scratch layout of launch_scatter_sorted: sorted keys | sorted values | run keys | run sums | run count | CUB temp storage
static size_t cub_tmp_bytes(uint64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const unsigned long long*)nullptr,
                                    (unsigned long long*)nullptr, (int64_t)n);
    cub::DeviceReduce::ReduceByKey(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const unsigned long long*)nullptr,
                                   (unsigned long long*)nullptr, (unsigned long long*)nullptr, cub::Sum(), (int64_t)n);
    return std::max(a, b);
}
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
size_t scatter_sorted_tmp_bytes(uint64_t n) { return 2 * al256(n * 4) + 2 * al256(n * 8) + 256 + al256(cub_tmp_bytes(n)) + 1024; }
void launch_scatter_sorted(const uint32_t* key, const unsigned long long* val, uint64_t n, unsigned long long* bases, void* tmp, size_t tmp_bytes,
                           cudaStream_t st) {
    if (n == 0) return;
    uint8_t* p = reinterpret_cast<uint8_t*>(tmp);
    uint32_t* skey = reinterpret_cast<uint32_t*>(p); p += al256(n * 4);
    unsigned long long* sval = reinterpret_cast<unsigned long long*>(p); p += al256(n * 8);
    uint32_t* rkey = reinterpret_cast<uint32_t*>(p); p += al256(n * 4);
    unsigned long long* rsum = reinterpret_cast<unsigned long long*>(p); p += al256(n * 8);
    unsigned long long* nrun = reinterpret_cast<unsigned long long*>(p); p += 256;
    size_t cb = tmp_bytes - (size_t)(p - reinterpret_cast<uint8_t*>(tmp));
    cub::DeviceRadixSort::SortPairs(p, cb, key, skey, val, sval, (int64_t)n, 0, 32, st);
    cb = tmp_bytes - (size_t)(p - reinterpret_cast<uint8_t*>(tmp));
    cub::DeviceReduce::ReduceByKey(p, cb, skey, rkey, sval, rsum, nrun, cub::Sum(), (int64_t)n, st);
    k_add_runs<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(rkey, rsum, nrun, bases);
    for (int i = 0; i < 3; ++i) PTX_LAUNCHED();
}
void launch_ds_rehash(const ulonglong2* old_slots, uint64_t old_cap, ulonglong2* new_slots, uint32_t new_shift, uint64_t new_mask, uint32_t ep,
                      cudaStream_t st) {
    k_ds_rehash<<<grid_for(old_cap, 256), 256, 0, st>>>(old_slots, old_cap, new_slots, new_shift, new_mask, ep);
    PTX_LAUNCHED();
}
void launch_ds_merge_boxes(const ulonglong2* inbox, const unsigned long long* off, const unsigned long long* cnt, uint32_t n_boxes, uint64_t max_cnt,
                           ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, uint32_t* flags, cudaStream_t st) {
    if (n_boxes == 0 || max_cnt == 0) return;
    dim3 grid(grid_for(max_cnt, 256, 148u * 8u), n_boxes);
    k_ds_merge_boxes<<<grid, 256, 0, st>>>(inbox, off, cnt, slots, shift, mask, ep, flags);
    PTX_LAUNCHED();
}
void launch_ds_collect_mixed(const ulonglong2* slots, uint64_t cap, uint32_t ep, unsigned long long* cursor, ulonglong2* out, uint64_t out_cap, cudaStream_t st) {
    k_ds_collect_mixed<<<grid_for(cap, 256), 256, 0, st>>>(slots, cap, ep, cursor, out, out_cap);
    PTX_LAUNCHED();
}
void launch_ds_apply_mixed(const ulonglong2* in, uint64_t n, ulonglong2* slots, uint32_t shift, uint64_t mask, uint32_t ep, uint32_t* scratch_flags,
                           cudaStream_t st) {
    if (n == 0) return;
    k_ds_apply_mixed<<<grid_for(n, 256), 256, 0, st>>>(in, n, slots, shift, mask, ep, scratch_flags);
    PTX_LAUNCHED();
}
void launch_flt_count_nl(const uint8_t* text, uint64_t n, uint32_t n_micro, uint32_t* cnt, cudaStream_t st) {
    k_flt_count_nl<<<(n_micro + 7) / 8, 256, 0, st>>>(text, n, n_micro, cnt);
    PTX_LAUNCHED();
}
void launch_flt_line_starts(const uint8_t* text, uint64_t n, uint32_t n_micro, const uint64_t* micro_base, uint64_t* line_off, cudaStream_t st) {
    k_flt_line_starts<<<(n_micro + 7) / 8, 256, 0, st>>>(text, n, n_micro, micro_base, line_off);
    PTX_LAUNCHED();
}
size_t flt_rec_bytes() { return sizeof(FltRec); }
void launch_flt_pipeline(const uint8_t* text, uint64_t n, const uint64_t* line_off, uint64_t n_lines, void* recs, ulonglong2* keys,
                         ulonglong2* best, unsigned long long* first, uint64_t mask, uint32_t shift, uint8_t* qual, uint32_t* sel,
                         uint32_t* flags, cudaStream_t st) {
    if (n_lines == 0) return;
    FltRec* r = reinterpret_cast<FltRec*>(recs);
    const uint32_t g128 = (uint32_t)((n_lines + 127) / 128), g256 = (uint32_t)((n_lines + 255) / 256);
    k_flt_parse<<<g128, 128, 0, st>>>(text, n, line_off, n_lines, r, flags);
    k_flt_best<<<g256, 256, 0, st>>>(r, n_lines, keys, best, mask, shift);
    k_flt_first<<<g256, 256, 0, st>>>(r, n_lines, keys, best, first, mask, shift, qual);
    k_flt_select<<<g256, 256, 0, st>>>(r, n_lines, keys, first, mask, shift, qual, sel);
    for (int i = 0; i < 4; ++i) PTX_LAUNCHED();
}
void launch_flt_compact(const uint32_t* sel, const uint64_t* scan, const uint64_t* line_off, uint64_t n_lines, uint64_t* out, uint64_t cap,
                        cudaStream_t st) {
    if (n_lines == 0) return;
    k_flt_compact<<<(uint32_t)((n_lines + 255) / 256), 256, 0, st>>>(sel, scan, line_off, n_lines, out, cap);
    PTX_LAUNCHED();
}
void launch_ninfo_build(const uint32_t* len, const uint64_t* bit_off, uint4* ninfo, int64_t N, cudaStream_t st) {
    if (N <= 0) return;
    k_ninfo_build<<<(uint32_t)((N + 255) / 256), 256, 0, st>>>(len, bit_off, ninfo, N);
    PTX_LAUNCHED();
}
void launch_ninfo_full(uint4* ninfo, uint8_t* full, int64_t N, int mode, cudaStream_t st) {
    if (N <= 0) return;
    k_ninfo_full<<<(uint32_t)((N + 255) / 256), 256, 0, st>>>(ninfo, full, N, mode);
    PTX_LAUNCHED();
}
void launch_mark_path_dups(uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint64_t* pbm_off, const uint32_t* pbase, uint32_t* bm,
                           cudaStream_t st) {
    if (P <= 0) return;
    k_mark_path_dups<<<grid_for((uint64_t)P, 256), 256, 0, st>>>(pnode, poff, Htot, P, pbm_off, pbase, bm);
    PTX_LAUNCHED();
}
void launch_path_len_sum(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint32_t* val, unsigned long long* out,
                         cudaStream_t st) {
    if (P <= 0) return;
    uint64_t nb = ((uint64_t)P + 256 * 16 - 1) / (256 * 16);
    k_path_len_sum<<<(uint32_t)nb, 256, 0, st>>>(pnode, poff, Htot, P, val, out);
    PTX_LAUNCHED();
}
void launch_trio_count(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, uint4* keys, uint32_t* cnt, uint32_t mask,
                       cudaStream_t st) {
    k_trio_count<<<grid_for(P, 256), 256, 0, st>>>(pnode, poff, Htot, P, keys, cnt, mask);
    PTX_LAUNCHED();
}
void launch_trio_flag(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint4* keys, const uint32_t* cnt,
                      uint32_t mask, uint32_t* flag, cudaStream_t st) {
    k_trio_flag<<<grid_for(P, 256), 256, 0, st>>>(pnode, poff, Htot, P, const_cast<uint4*>(keys), cnt, mask, flag);
    PTX_LAUNCHED();
}
void launch_scan_u32(const uint32_t* in, uint64_t* out, uint64_t n, uint64_t* scratch, cudaStream_t st) {
    const uint64_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    k_scan_block_sums<<<(uint32_t)std::max<uint64_t>(nb, 1), 1024, 0, st>>>(in, n, scratch);
    k_scan_bsums<<<1, 1024, 0, st>>>(scratch, std::max<uint64_t>(nb, 1));
    k_scan_final<<<(uint32_t)std::max<uint64_t>(nb, 1), 1024, 0, st>>>(in, n, scratch, std::max<uint64_t>(nb, 1), out);
    PTX_LAUNCHED();
    PTX_LAUNCHED();
    PTX_LAUNCHED();
}
void launch_trio_emit(const uint32_t* pnode, const uint64_t* poff, int64_t Htot, int64_t P, const uint32_t* flag, const uint64_t* scan,
                      const uint32_t* len, uint32_t* trio_key, int64_t* trio_len, uint32_t* trio_owner, uint4* tt, uint32_t tt_mask,
                      uint4* ninfo, uint64_t* trio_start, cudaStream_t st) {
    k_trio_emit<<<grid_for(std::max<int64_t>(P, Htot + 1), 256), 256, 0, st>>>(pnode, poff, Htot, P, flag, scan, len, trio_key, trio_len,
                                                                               trio_owner, tt, tt_mask, ninfo, trio_start);
    PTX_LAUNCHED();
}
void launch_cov(const GraphDev& g, bool bits_only, cudaStream_t st) {
    if (g.N <= 0) return;
    k_cov<<<(uint32_t)((g.N + 255) / 256), 256, 0, st>>>(g.len, g.bit_off, bits_only ? nullptr : g.full, g.bits, g.cov, g.N);
    PTX_LAUNCHED();
}
void launch_path_cov_sum(const GraphDev& g, cudaStream_t st) {
    launch_path_len_sum(g.pnode, g.poff, g.Htot, g.P, g.cov, g.path_cov_sum, st);
}
void launch_hap_nz(const GraphDev& g, cudaStream_t st) {
    if (g.T <= 0) return;
    k_hap_nz<<<(uint32_t)((g.T + 255) / 256), 256, 0, st>>>(g.trio_bases, g.trio_owner, g.T, g.hap_nz);
    PTX_LAUNCHED();
}
void launch_depth(const unsigned long long* num, const uint32_t* den32, const int64_t* den64, double* out, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    k_depth<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(num, den32, den64, out, n);
    PTX_LAUNCHED();
}
void launch_bits_fill_full(const GraphDev& g, cudaStream_t st) {
    if (g.N <= 0) return;
    k_bits_fill_full<<<(uint32_t)((g.N + 255) / 256), 256, 0, st>>>(g.ninfo, g.bits, g.N);
    PTX_LAUNCHED();
}
void launch_or_slices(uint32_t* dst, const uint32_t* src, uint32_t n_src, uint64_t n_words, cudaStream_t st) {
    if (n_words == 0 || n_src == 0) return;
    k_or_slices<<<grid_for(n_words, 256 * 4), 256, 0, st>>>(dst, src, n_src, n_words);
    PTX_LAUNCHED();
}
void launch_or_words(uint32_t* dst, const uint32_t* src, uint64_t n_words, cudaStream_t st) {
    if (n_words == 0) return;
    k_or_words<<<grid_for(n_words, 256 * 4), 256, 0, st>>>(dst, src, n_words);
    PTX_LAUNCHED();
}

}  // namespace ptx
