//! Raw bindings of include/pantax_gpu.h + a small safe wrapper used by the patch in INTEGRATION.md.
//! SOURCE ONLY - never compiled in the build image (no rustc there).
#![allow(non_camel_case_types)]
use libc::{c_char, c_double, c_int, c_void, size_t};

#[repr(C)]
pub struct ptx_ctx { _private: [u8; 0] }

pub const PTX_OK: c_int = 0;
pub const PTX_E_START_GT_LEN: c_int = -7;
pub const PTX_LABEL_UNCLASSIFIED: u32 = 0xFFFF_FFFF;

extern "C" {
    pub fn ptx_create(device: c_int, out: *mut *mut ptx_ctx) -> c_int;
    pub fn ptx_destroy(ctx: *mut ptx_ctx);
    pub fn ptx_last_error(ctx: *const ptx_ctx) -> *const c_char;
    pub fn ptx_version() -> *const c_char;
    pub fn ptx_set_ranges(ctx: *mut ptx_ctx, n: c_int, taxid: *const *const c_char, start: *const i64, end: *const i64) -> c_int;
    pub fn ptx_upload_graph(ctx: *mut ptx_ctx, species: c_int, nodes_len: *const i64, n: i64, path_off: *const u64,
                            path_nodes: *const u64, n_paths: i64) -> c_int;
    pub fn ptx_upload_graph_gfa(ctx: *mut ptx_ctx, species: c_int, gfa: *const u8, n: usize) -> c_int;
    pub fn ptx_species_graph(ctx: *mut ptx_ctx, species: c_int, nodes_len: *mut i64, path_off: *mut u64, path_nodes: *mut u64) -> c_int;
    pub fn ptx_species_path_steps(ctx: *const ptx_ctx, species: c_int) -> i64;
    pub fn ptx_species_path_name(ctx: *mut ptx_ctx, species: c_int, h: i64, buf: *mut c_char, cap: usize) -> c_int;
    pub fn ptx_commit_graphs(ctx: *mut ptx_ctx) -> c_int;
    pub fn ptx_reserve(ctx: *mut ptx_ctx, expected_records: i64) -> c_int;
    pub fn ptx_host_alloc(bytes: size_t, out: *mut *mut c_void) -> c_int;
    pub fn ptx_host_free(p: *mut c_void) -> c_int;
    pub fn ptx_ingest_gaf(ctx: *mut ptx_ctx, bytes: *const u8, n: size_t, is_last: c_int) -> c_int;
    pub fn ptx_gaf_buffer_alloc(ctx: *mut ptx_ctx, capacity: size_t, buffer_id: *mut c_int, device_ptr: *mut *mut c_void) -> c_int;
    pub fn ptx_ingest_gaf_device(ctx: *mut ptx_ctx, buffer_id: c_int, n: size_t) -> c_int;
    pub fn ptx_ingest_labels(ctx: *mut ptx_ctx, labels: *const u32, n: i64) -> c_int;
    pub fn ptx_finalize(ctx: *mut ptx_ctx) -> c_int;
    pub fn ptx_reset(ctx: *mut ptx_ctx) -> c_int;
    pub fn ptx_rewind(ctx: *mut ptx_ctx) -> c_int;
    pub fn ptx_num_records(ctx: *const ptx_ctx) -> i64;
    pub fn ptx_num_species(ctx: *const ptx_ctx) -> c_int;
    pub fn ptx_ids_unique(ctx: *const ptx_ctx) -> c_int;
    pub fn ptx_read_labels(ctx: *mut ptx_ctx, labels: *mut u32) -> c_int;
    pub fn ptx_species_counts(ctx: *mut ptx_ctx, counts: *mut i64) -> c_int;
    pub fn ptx_equal_length(ctx: *mut ptx_ctx, is_equal: *mut c_int, read_len: *mut i64) -> c_int;
    pub fn ptx_species_nodes(ctx: *const ptx_ctx, species: c_int) -> i64;
    pub fn ptx_species_paths(ctx: *const ptx_ctx, species: c_int) -> i64;
    pub fn ptx_species_trios(ctx: *const ptx_ctx, species: c_int) -> i64;
    pub fn ptx_node_bases(ctx: *mut ptx_ctx, species: c_int, bases: *mut i64) -> c_int;
    pub fn ptx_node_cov(ctx: *mut ptx_ctx, species: c_int, cov: *mut u64) -> c_int;
    pub fn ptx_node_depth(ctx: *mut ptx_ctx, species: c_int, depth: *mut c_double) -> c_int;
    pub fn ptx_trio_bases(ctx: *mut ptx_ctx, species: c_int, bases: *mut i64) -> c_int;
    pub fn ptx_trio_depth(ctx: *mut ptx_ctx, species: c_int, depth: *mut c_double) -> c_int;
    pub fn ptx_trio_table(ctx: *mut ptx_ctx, species: c_int, keys3: *mut u64, len: *mut i64, owner: *mut u32) -> c_int;
    pub fn ptx_trio_ref_order(path_off: *const u64, path_nodes: *const u64, n_paths: i64, keys3: *const u64, n_trios: i64, order: *mut u64) -> c_int;
    pub fn ptx_path_sums(ctx: *mut ptx_ctx, species: c_int, sum_cov: *mut i64, sum_len: *mut i64) -> c_int;
    pub fn ptx_hap_trio_counts(ctx: *mut ptx_ctx, species: c_int, u: *mut i64, nz: *mut i64) -> c_int;
    pub fn ptx_filter_gaf(ctx: *mut ptx_ctx, bytes: *const u8, n: size_t, out_line_off: *mut u64, cap: i64, n_out: *mut i64) -> c_int;
    pub fn ptx_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn ptx_comm_init(ctx: *mut ptx_ctx, n_ranks: c_int, rank: c_int, id128: *const c_void) -> c_int;
    pub fn ptx_create_multi(devices: *const c_int, n_devices: c_int, expected_records_per_device: i64, out: *mut *mut ptx_ctx) -> c_int;
    pub fn ptx_finalize_multi(ctxs: *const *mut ptx_ctx, n: c_int) -> c_int;
    pub fn ptx_stats_json(ctx: *mut ptx_ctx, buf: *mut c_char, cap: size_t) -> c_int;
    pub fn ptx_timing(ctx: *mut ptx_ctx, ingest_ms: *mut c_double, finalize_ms: *mut c_double, kernel_launches: *mut i64) -> c_int;
}

/// Safe handle.  One per process per GPU; not `Sync` (a ctx is driven by one host thread).
pub struct Gpu { raw: *mut ptx_ctx }

#[derive(Debug)]
pub struct GpuError { pub code: i32, pub msg: String }

/// ptx_finalize of every context of a `Gpu::new_multi` group (the collectives inside need all of them at once).
pub fn finalize_multi(gpus: &[Gpu]) -> Result<(), GpuError> {
    let raw: Vec<*mut ptx_ctx> = gpus.iter().map(|g| g.raw).collect();
    let rc = unsafe { ptx_finalize_multi(raw.as_ptr(), raw.len() as c_int) };
    if rc == 0 { Ok(()) } else { Err(GpuError { code: rc, msg: gpus.iter().map(|g| g.last_error()).collect::<Vec<_>>().join("; ") }) }
}

impl Gpu {
    /// One context per device, all in this process (the `pantax` CLI is one process), joined into one communicator
    /// group: the 8-GPU form of `Gpu::new(0)`.  `strain_profiling` (profile.rs:3291) feeds context k the k-th read
    /// batch of the GAF, calls `finalize_multi`, and reads the (identical, reduced) results from any of them.
    pub fn new_multi(devices: &[i32], expected_records_per_device: usize) -> Result<Vec<Gpu>, GpuError> {
        let mut raw: Vec<*mut ptx_ctx> = vec![std::ptr::null_mut(); devices.len()];
        let rc = unsafe { ptx_create_multi(devices.as_ptr(), devices.len() as c_int, expected_records_per_device as i64, raw.as_mut_ptr()) };
        if rc != 0 { return Err(GpuError { code: rc, msg: "ptx_create_multi".into() }); }
        Ok(raw.into_iter().map(|r| Gpu { raw: r }).collect())
    }
    pub fn new(device: i32) -> Result<Gpu, GpuError> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { ptx_create(device, &mut raw) };
        if rc != PTX_OK { return Err(GpuError { code: rc, msg: "ptx_create: no CUDA device (no CPU fallback)".into() }); }
        Ok(Gpu { raw })
    }
    pub fn last_error(&self) -> String {
        unsafe { std::ffi::CStr::from_ptr(ptx_last_error(self.raw)) }.to_string_lossy().into_owned()
    }
    fn ck(&self, rc: c_int) -> Result<(), GpuError> {
        if rc == PTX_OK { return Ok(()); }
        let msg = unsafe { std::ffi::CStr::from_ptr(ptx_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(GpuError { code: rc, msg })
    }
    /// species_range.txt rows in file order (rcls.rs:40-71).
    pub fn set_ranges(&self, ranges: &[(String, i64, i64)]) -> Result<(), GpuError> {
        let names: Vec<std::ffi::CString> = ranges.iter().map(|r| std::ffi::CString::new(r.0.as_str()).unwrap()).collect();
        let ptrs: Vec<*const c_char> = names.iter().map(|s| s.as_ptr()).collect();
        let st: Vec<i64> = ranges.iter().map(|r| r.1).collect();
        let en: Vec<i64> = ranges.iter().map(|r| r.2).collect();
        self.ck(unsafe { ptx_set_ranges(self.raw, ranges.len() as c_int, ptrs.as_ptr(), st.as_ptr(), en.as_ptr()) })
    }
    /// types.rs:51-55 `Graph` of one species; BTreeMap iteration gives the hap-name order the ABI wants.
    pub fn upload_graph(&self, species: usize, nodes_len: &[i64], paths: &std::collections::BTreeMap<String, Vec<usize>>) -> Result<(), GpuError> {
        let mut off = vec![0u64];
        let mut flat: Vec<u64> = Vec::new();
        for p in paths.values() { flat.extend(p.iter().map(|&v| v as u64)); off.push(flat.len() as u64); }
        if flat.is_empty() { flat.push(0); }
        self.ck(unsafe { ptx_upload_graph(self.raw, species as c_int, nodes_len.as_ptr(), nodes_len.len() as i64, off.as_ptr(), flat.as_ptr(), paths.len() as i64) })
    }
    /// profile.rs:466-545 `read_gfa(&gfa_file, 0)` on the device: hand over the file's bytes instead of a parsed `Graph`.
    pub fn upload_graph_gfa(&self, species: usize, gfa: &[u8]) -> Result<(), GpuError> {
        self.ck(unsafe { ptx_upload_graph_gfa(self.raw, species as c_int, gfa.as_ptr(), gfa.len()) })
    }
    pub fn commit_graphs(&self) -> Result<(), GpuError> { self.ck(unsafe { ptx_commit_graphs(self.raw) }) }
    pub fn ingest_gaf(&self, bytes: &[u8], is_last: bool) -> Result<(), GpuError> {
        self.ck(unsafe { ptx_ingest_gaf(self.raw, bytes.as_ptr(), bytes.len(), is_last as c_int) })
    }
    /// Strain-only resume (profile.rs:3365-3385): species of every GAF row from reads_classification.tsv
    /// (index into the ranges, u32::MAX = "U"); call before ingest_gaf.
    pub fn ingest_labels(&self, labels: &[u32]) -> Result<(), GpuError> {
        self.ck(unsafe { ptx_ingest_labels(self.raw, labels.as_ptr(), labels.len() as i64) })
    }
    /// Expected number of GAF records; before comm_init it also sizes the peer-memory id boxes.
    pub fn reserve(&self, records: usize) -> Result<(), GpuError> { self.ck(unsafe { ptx_reserve(self.raw, records as i64) }) }
    pub fn finalize(&self) -> Result<(), GpuError> { self.ck(unsafe { ptx_finalize(self.raw) }) }
    /// profile.rs:1112-1135: (unique trios owned, of those with depth > 0) per hap.
    pub fn hap_trio_counts(&self, species: usize) -> Result<(Vec<i64>, Vec<i64>), GpuError> {
        let s = species as c_int;
        let h = unsafe { ptx_species_paths(self.raw, s) }.max(0) as usize;
        let (mut a, mut b) = (vec![0i64; h.max(1)], vec![0i64; h.max(1)]);
        self.ck(unsafe { ptx_hap_trio_counts(self.raw, s, a.as_mut_ptr(), b.as_mut_ptr()) })?;
        a.truncate(h); b.truncate(h);
        Ok((a, b))
    }
    pub fn num_records(&self) -> usize { unsafe { ptx_num_records(self.raw) as usize } }
    pub fn read_labels(&self) -> Result<Vec<u32>, GpuError> {
        let mut v = vec![0u32; self.num_records().max(1)];
        self.ck(unsafe { ptx_read_labels(self.raw, v.as_mut_ptr()) })?;
        v.truncate(self.num_records());
        Ok(v)
    }
    pub fn species_counts(&self, n_species: usize) -> Result<Vec<[i64; 4]>, GpuError> {
        let mut v = vec![[0i64; 4]; n_species];
        self.ck(unsafe { ptx_species_counts(self.raw, v.as_mut_ptr() as *mut i64) })?;
        Ok(v)
    }
    /// get_node_abundances (profile.rs:743-1026): (node_abundance_vec, trio_node_abundance_vec, node_base_cov).
    pub fn get_node_abundances(&self, species: usize) -> Result<(Vec<f64>, Vec<f64>, Vec<usize>), GpuError> {
        let s = species as c_int;
        let n = unsafe { ptx_species_nodes(self.raw, s) }.max(0) as usize;
        let t = unsafe { ptx_species_trios(self.raw, s) }.max(0) as usize;
        let (mut depth, mut tdepth, mut cov) = (vec![0f64; n.max(1)], vec![0f64; t.max(1)], vec![0u64; n.max(1)]);
        self.ck(unsafe { ptx_node_depth(self.raw, s, depth.as_mut_ptr()) })?;
        self.ck(unsafe { ptx_trio_depth(self.raw, s, tdepth.as_mut_ptr()) })?;
        self.ck(unsafe { ptx_node_cov(self.raw, s, cov.as_mut_ptr()) })?;
        depth.truncate(n); tdepth.truncate(t); cov.truncate(n);
        Ok((depth, tdepth, cov.into_iter().map(|x| x as usize).collect()))
    }
    /// gaf_filter::filter_max_alignment_mt (gaf_filter.rs:44-97) on a GAF held in memory: byte offsets of the kept lines, file order.
    /// alignment.rs:171 becomes: read the file, call this, write `&bytes[off..line_end]` + '\n' per offset to `<stem>_filtered.gaf`.
    pub fn filter_gaf(&self, bytes: &[u8]) -> Result<Vec<u64>, GpuError> {
        let cap = bytes.iter().filter(|&&c| c == b'\n').count() + 1;
        let mut off = vec![0u64; cap];
        let mut n_out: i64 = 0;
        self.ck(unsafe { ptx_filter_gaf(self.raw, bytes.as_ptr(), bytes.len(), off.as_mut_ptr(), cap as i64, &mut n_out) })?;
        off.truncate(n_out as usize);
        Ok(off)
    }
    /// (sum of node_base_cov, sum of nodes_len) over the distinct nodes of every path (profile.rs:2714-2724).
    pub fn path_sums(&self, species: usize) -> Result<(Vec<i64>, Vec<i64>), GpuError> {
        let s = species as c_int;
        let h = unsafe { ptx_species_paths(self.raw, s) }.max(0) as usize;
        let (mut a, mut b) = (vec![0i64; h.max(1)], vec![0i64; h.max(1)]);
        self.ck(unsafe { ptx_path_sums(self.raw, s, a.as_mut_ptr(), b.as_mut_ptr()) })?;
        a.truncate(h); b.truncate(h);
        Ok((a, b))
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { ptx_destroy(self.raw) } } }

/// order[i] = row of `keys3` (ptx_trio_table, `n_trios` x 3 canonical local ids) that `trio_nodes_info` numbers i
/// (profile.rs:659-716: FxHashSet iteration order restricted to the trios that occur once).  Inside the reference crate the
/// real FxHashSet is at hand and this is not needed; it exists for callers that want the reference's f64 summation
/// order of `frequencies_mean` (profile.rs:1123-1146) without building that set.  `None`: the paths and the table disagree.
pub fn trio_ref_order(paths: &std::collections::BTreeMap<String, Vec<usize>>, keys3: &[u64]) -> Option<Vec<usize>> {
    let mut off = vec![0u64];
    let mut flat: Vec<u64> = Vec::new();
    for p in paths.values() { flat.extend(p.iter().map(|&v| v as u64)); off.push(flat.len() as u64); }
    if flat.is_empty() { flat.push(0); }
    let t = keys3.len() / 3;
    let mut order = vec![0u64; t.max(1)];
    let rc = unsafe { ptx_trio_ref_order(off.as_ptr(), flat.as_ptr(), paths.len() as i64, keys3.as_ptr(), t as i64, order.as_mut_ptr()) };
    if rc != 0 { return None; }
    order.truncate(t);
    Some(order.into_iter().map(|x| x as usize).collect())
}

/// profile.rs:2714-2729: the reference forms `path_cov_ratio` as `RowDVector<f32>(node_base_cov) * incidence` over
/// `RowDVector<f32>(node_len) * incidence`; nalgebra evaluates each product as a gemv, i.e. a sequential f32
/// accumulation over the nodes in index order (nodes outside the path contribute 0).  This repeats exactly that over
/// the distinct nodes of `path`, so the value is the reference's even where the running sums pass 2^24.
pub fn path_cov_ratio_f32(path: &[usize], node_base_cov: &[usize], node_len: &[i64]) -> f32 {
    let mut on_path = vec![false; node_len.len()];
    for &v in path { on_path[v] = true; }
    let (mut cov, mut len) = (0f32, 0f32);
    for v in 0..node_len.len() {
        if on_path[v] { cov += node_base_cov[v] as f32; len += node_len[v] as f32; }
    }
    cov / len
}
/// The f64 twin used by cbc_opt (profile.rs:1952-1977); exact below 2^53.
pub fn path_cov_ratio_f64(path: &[usize], node_base_cov: &[usize], node_len: &[i64]) -> f64 {
    let mut on_path = vec![false; node_len.len()];
    for &v in path { on_path[v] = true; }
    let (mut cov, mut len) = (0f64, 0f64);
    for v in 0..node_len.len() {
        if on_path[v] { cov += node_base_cov[v] as f64; len += node_len[v] as f64; }
    }
    cov / len
}
