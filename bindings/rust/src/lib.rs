//! bindings/rust/src/lib.rs — Rust side of the drop-in boundary (include/vmis.h).
//!
//! `GpuVMISIndex` takes the place of `VMISIndex` (reference src/vmisknn/vmis_index.rs:28-35) and implements the
//! reference's plugin trait `SimilarityComputationNew` (src/vmisknn/similarity_indexed.rs:8-24); `predict` has the
//! signature of `vmisknn::predict` (src/vmisknn/mod.rs:118-125).  Inside bolcom/serenade this file is
//! `src/vmisknn/gpu_index.rs` and the three `use` lines below point at the crate's own types; the stand-alone
//! definitions under `mod reference_types` mirror them so that the file is self-contained.
//!
//! Not compiled in this repository (no Rust toolchain in the build image); every `extern "C"` item is checked
//! against the exported symbols of libvmis_b200.so by tests/test_host.py::test_rust_binding_matches_header.
use std::collections::BinaryHeap;
use std::ffi::{c_void, CStr, CString};
use std::os::raw::{c_char, c_double, c_int};

pub use reference_types::{ItemScore, ProductAttributes, SessionScore, SimilarityComputationNew};

/// The reference's types on this path, restated (mod.rs:15-74, vmis_index.rs:23-26, similarity_indexed.rs:8-24).
pub mod reference_types {
    use std::cmp::Ordering;
    use std::collections::BinaryHeap;

    #[derive(Debug, Clone, PartialEq)]
    pub struct ProductAttributes { pub is_adult: bool, pub is_for_sale: bool }

    #[derive(PartialEq, Debug)]
    pub struct SessionScore { pub id: u32, pub score: f64 }
    #[derive(PartialEq, Debug)]
    pub struct ItemScore { pub id: u64, pub score: f64 }

    // min-heap order on the score (mod.rs:24-43, :54-74): `peek()` is the eviction candidate and
    // `into_sorted_vec()` is score-descending.
    impl Eq for SessionScore {}
    impl Ord for SessionScore {
        fn cmp(&self, other: &Self) -> Ordering { other.score.partial_cmp(&self.score).unwrap_or(Ordering::Equal) }
    }
    impl PartialOrd for SessionScore { fn partial_cmp(&self, other: &Self) -> Option<Ordering> { Some(self.cmp(other)) } }
    impl Eq for ItemScore {}
    impl Ord for ItemScore {
        fn cmp(&self, other: &Self) -> Ordering { other.score.partial_cmp(&self.score).unwrap_or(Ordering::Equal) }
    }
    impl PartialOrd for ItemScore { fn partial_cmp(&self, other: &Self) -> Option<Ordering> { Some(self.cmp(other)) } }

    pub trait SimilarityComputationNew {
        fn items_for_session(&self, session: &u32) -> &[u64];
        fn idf(&self, item_id: &u64) -> f64;
        fn find_neighbors(&self, evolving_session: &[u64], k: usize, m: usize) -> BinaryHeap<SessionScore>;
        fn find_attributes(&self, item_id: &u64) -> Option<&ProductAttributes>;
    }
}

#[repr(C)]
pub struct vmis_index_t { _private: [u8; 0] }

pub const VMIS_ATTR_EXISTS: c_int = 1;
pub const VMIS_ATTR_FOR_SALE: c_int = 2;
pub const VMIS_ATTR_ADULT: c_int = 4;

extern "C" {
    // include/vmis.h — constructors
    pub fn vmis_index_from_csv(path: *const c_char, m: usize, idf_weighting: c_double, device: c_int) -> *mut vmis_index_t;
    pub fn vmis_index_from_avro(base_path: *const c_char, device: c_int) -> *mut vmis_index_t;
    pub fn vmis_index_from_sessions_attrs(items: *const u64, sess_off: *const u64, sess_ts: *const u32, n_sessions: usize,
                                          m: usize, max_len: usize, idf_weighting: c_double, attr_items: *const u64,
                                          attr_flags: *const u8, n_attrs: usize, device: c_int) -> *mut vmis_index_t;
    pub fn vmis_index_free(index: *mut vmis_index_t);
    // queries
    pub fn vmis_predict(index: *const vmis_index_t, evolving_session: *const u64, len: usize, k: usize, m: usize,
                        how_many: usize, enable_business_logic: c_int, out_ids: *mut u64, out_scores: *mut f64) -> c_int;
    pub fn vmis_predict_batch(index: *const vmis_index_t, q_items: *const u64, q_off: *const u32, n_q: u32, k: u32,
                              m: u32, how_many: u32, enable_business_logic: c_int, out_ids: *mut u64,
                              out_scores: *mut f64, out_counts: *mut u32, stream: *mut c_void) -> c_int;
    pub fn vmis_find_neighbors_batch(index: *const vmis_index_t, q_items: *const u64, q_off: *const u32, n_q: u32,
                                     k: u32, m: u32, out_sess: *mut u32, out_sim: *mut f64, out_counts: *mut u32,
                                     stream: *mut c_void) -> c_int;
    // trait accessors (host mirror)
    pub fn vmis_items_for_session(index: *const vmis_index_t, session: u32, len: *mut usize) -> *const u64;
    pub fn vmis_idf(index: *const vmis_index_t, item: u64, out: *mut f64) -> c_int;
    pub fn vmis_find_attributes(index: *const vmis_index_t, item: u64) -> c_int;
    pub fn vmis_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(vmis_last_error()) }.to_string_lossy().into_owned()
}

/// HBM-resident replacement of `VMISIndex`.
pub struct GpuVMISIndex {
    h: *mut vmis_index_t,
    // `find_attributes` hands out references (similarity_indexed.rs:23); the four possible attribute values live here
    // so that a reference can be returned without a per-item map on the Rust side.  Index = adult | for_sale << 1.
    attr_values: [ProductAttributes; 4],
}
unsafe impl Send for GpuVMISIndex {} // the C ABI is re-entrant (include/vmis.h)
unsafe impl Sync for GpuVMISIndex {}

impl GpuVMISIndex {
    fn wrap(h: *mut vmis_index_t) -> Self {
        if h.is_null() {
            panic!("{}", last_error()); // the reference's builders unwrap() their I/O (vmis_index.rs:50)
        }
        let pa = |adult, sale| ProductAttributes { is_adult: adult, is_for_sale: sale };
        GpuVMISIndex { h, attr_values: [pa(false, false), pa(true, false), pa(false, true), pa(true, true)] }
    }
    /// same arguments as `VMISIndex::new_from_csv` (vmis_index.rs:38)
    pub fn new_from_csv(path_to_training: &str, m_most_recent_sessions: usize, idf_weighting: f64) -> Self {
        let c = CString::new(path_to_training).unwrap();
        Self::wrap(unsafe { vmis_index_from_csv(c.as_ptr(), m_most_recent_sessions, idf_weighting, 0) })
    }
    /// same argument as `VMISIndex::new` (vmis_index.rs:85): `<base_path>/itemindex/*.avro` + `sessionindex/*.avro`
    pub fn new(base_path: &str) -> Self {
        let c = CString::new(base_path).unwrap();
        Self::wrap(unsafe { vmis_index_from_avro(c.as_ptr(), 0) })
    }
    pub fn raw(&self) -> *const vmis_index_t { self.h }
}
impl Drop for GpuVMISIndex {
    fn drop(&mut self) { unsafe { vmis_index_free(self.h) } }
}

impl SimilarityComputationNew for GpuVMISIndex {
    fn items_for_session(&self, session: &u32) -> &[u64] {
        let mut n = 0usize;
        let p = unsafe { vmis_items_for_session(self.h, *session, &mut n) };
        if p.is_null() { panic!("{}", last_error()); } // vmis_index.rs:318 indexes out of bounds → panic
        unsafe { std::slice::from_raw_parts(p, n) }
    }
    fn idf(&self, item_id: &u64) -> f64 {
        let mut v = 0f64;
        if unsafe { vmis_idf(self.h, *item_id, &mut v) } != 0 { panic!("{}", last_error()) } // vmis_index.rs:322
        v
    }
    fn find_neighbors(&self, evolving_session: &[u64], k: usize, m: usize) -> BinaryHeap<SessionScore> {
        let off = [0u32, evolving_session.len() as u32];
        let (mut s, mut sim, mut cnt) = (vec![0u32; k.max(1)], vec![0f64; k.max(1)], 0u32);
        let rc = unsafe {
            vmis_find_neighbors_batch(self.h, evolving_session.as_ptr(), off.as_ptr(), 1, k as u32, m as u32,
                                      s.as_mut_ptr(), sim.as_mut_ptr(), &mut cnt, std::ptr::null_mut())
        };
        if rc != 0 { panic!("{}", last_error()) }
        (0..cnt as usize).map(|i| SessionScore { id: s[i], score: sim[i] }).collect()
    }
    /// bridges to `vmis_find_attributes` (vmis_index.rs:417-419): the attributes live in the handle — ForSale / IsAdult
    /// of the Avro item index, or `{adult:false, for_sale:true}` for a CSV-built index (vmis_index.rs:514-518)
    fn find_attributes(&self, item_id: &u64) -> Option<&ProductAttributes> {
        let a = unsafe { vmis_find_attributes(self.h, *item_id) };
        if a & VMIS_ATTR_EXISTS == 0 { return None; }
        let idx = ((a & VMIS_ATTR_ADULT != 0) as usize) | (((a & VMIS_ATTR_FOR_SALE != 0) as usize) << 1);
        Some(&self.attr_values[idx])
    }
}

/// Same signature as `vmisknn::predict` (mod.rs:118-125).  The accessors of the trait are only ever called from inside
/// `predict` (mod.rs:131,145,186,192,203), so replacing `predict` wholesale is the drop-in.  `into_sorted_vec()` on
/// the returned heap is score-descending, as at the call sites (recommend_resource.rs:58-62).
pub fn predict(index: &GpuVMISIndex, evolving_session: &[u64], k: usize, m: usize, how_many: usize,
               enable_business_logic: bool) -> BinaryHeap<ItemScore> {
    let (mut ids, mut sc) = (vec![0u64; how_many.max(1)], vec![0f64; how_many.max(1)]);
    let n = unsafe {
        vmis_predict(index.h, evolving_session.as_ptr(), evolving_session.len(), k, m, how_many,
                     enable_business_logic as c_int, ids.as_mut_ptr(), sc.as_mut_ptr())
    };
    if n < 0 { panic!("{}", last_error()) }
    (0..n as usize).map(|i| ItemScore { id: ids[i], score: sc[i] }).collect()
}

/// The batched shape of the evaluator / HPO objective (evaluator.rs:46-76, objective.rs:20-47): the whole replay in
/// one call.  Returns one `Vec<ItemScore>` per evolving session, best first.
pub fn predict_batch(index: &GpuVMISIndex, sessions: &[Vec<u64>], k: usize, m: usize, how_many: usize,
                     enable_business_logic: bool) -> Vec<Vec<ItemScore>> {
    let mut q_items = Vec::new();
    let mut q_off = vec![0u32];
    for s in sessions { q_items.extend_from_slice(s); q_off.push(q_items.len() as u32); }
    let n_q = sessions.len();
    let (mut ids, mut sc, mut cnt) = (vec![0u64; n_q * how_many.max(1)], vec![0f64; n_q * how_many.max(1)], vec![0u32; n_q]);
    let rc = unsafe {
        vmis_predict_batch(index.h, q_items.as_ptr(), q_off.as_ptr(), n_q as u32, k as u32, m as u32, how_many as u32,
                           enable_business_logic as c_int, ids.as_mut_ptr(), sc.as_mut_ptr(), cnt.as_mut_ptr(),
                           std::ptr::null_mut())
    };
    if rc != 0 { panic!("{}", last_error()) }
    (0..n_q).map(|q| (0..cnt[q] as usize).map(|i| ItemScore { id: ids[q * how_many + i], score: sc[q * how_many + i] }).collect()).collect()
}
