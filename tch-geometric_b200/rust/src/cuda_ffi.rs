//! `extern "C"` declarations of include/tchgeo_cuda.h for the reference crate (new file src/cuda_ffi.rs).
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type tchgeo_status = i32;
pub const TCHGEO_OK: tchgeo_status = 0;
pub const TCHGEO_SAMPLER_UNIFORM: i32 = 0;
pub const TCHGEO_SAMPLER_UNIFORM_REPLACE: i32 = 1;
pub const TCHGEO_SAMPLER_WEIGHTED: i32 = 2;

#[repr(C)]
pub struct tchgeo_sampling_args {
    pub num_node_types: i32, pub num_rels: i32, pub num_hops: i32, pub sampler_kind: i32,
    pub rel_src: *const i32, pub rel_dst: *const i32,
    pub col_ptrs: *const *const i64, pub num_cols: *const i64, pub row_indices: *const *const i64,
    pub weights: *const *const f64, pub row_indices32: *const *const i32,
    pub weights_cumsum: *const *const f64,
    pub fanouts: *const i64, pub rel_active: *const u8,
    pub num_batches: i64, pub inputs: *const *const i64, pub seeds_per_batch: *const i64,
    pub seed: u64, pub batch_base: u32, pub reserved0: u32,
    pub samples: *const *mut i64, pub samples_stride: *const i64,
    pub rows: *const *mut i64, pub cols: *const *mut i64, pub edge_index: *const *mut i64,
    pub edges_stride: *const i64,
    pub samples_len: *mut i64, pub edges_len: *mut i64, pub layer_offsets: *mut i64,
    pub filter_mode: i32, pub filter_forward: i32, pub filter_window_lo: i64, pub filter_window_hi: i64,
    pub timestamps: *const *const i64, pub inputs_state: *const *const i64, pub states: *const *mut i64,
    pub workspace: *mut c_void, pub workspace_bytes: usize, pub stream: *mut c_void,
}

/// struct tchgeo_negative_args (negative_sample_neighbors_*, src/python.rs:689-783)
#[repr(C)]
pub struct tchgeo_negative_args {
    pub num_node_types: i32, pub num_rels: i32,
    pub rel_src: *const i32, pub rel_dst: *const i32,
    pub row_ptrs: *const *const i64, pub col_indices: *const *const i64,
    pub num_rows: *const i64, pub node_count: *const i64,
    pub inputs: *const *const i64, pub num_inputs: *const i64,
    pub num_neg: i64, pub try_count: i64, pub inbound: i32, pub reserved0: i32, pub seed: u64,
    pub samples: *const *mut i64, pub rows: *const *mut i64, pub cols: *const *mut i64,
    pub samples_len: *mut i64, pub edges_len: *mut i64,
    pub workspace: *mut c_void, pub workspace_bytes: usize, pub stream: *mut c_void,
}

extern "C" {
    pub fn tchgeo_abi_version() -> i32;
    pub fn tchgeo_last_error() -> *const c_char;
    pub fn tchgeo_coo_to_csx_workspace_bytes(num_edges: i64, n_rows: i64, n_cols: i64) -> usize;
    pub fn tchgeo_coo_to_csx(row: *const i64, col: *const i64, num_edges: i64, n_rows: i64, n_cols: i64, csc: i32,
                             ptrs: *mut i64, indices: *mut i64, perm: *mut i64,
                             workspace: *mut c_void, workspace_bytes: usize, stream: *mut c_void) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_capacity(args: *const tchgeo_sampling_args, samples_cap: *mut i64,
                                             edges_cap: *mut i64) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_workspace_bytes(args: *const tchgeo_sampling_args) -> usize;
    pub fn tchgeo_neighbor_sampling(args: *const tchgeo_sampling_args) -> tchgeo_status;
    pub fn tchgeo_random_walk(row_ptrs: *const i64, num_rows: i64, col_indices: *const i64, start: *const i64,
                              num_walks: i64, walk_length: i64, p: f32, q: f32, seed: u64, walker_base: i64,
                              walks: *mut i64, stats: *mut i64, attempts_out: *mut i64,
                              stream: *mut c_void) -> tchgeo_status;
    // temporal walk, negative sampling, per-column transforms, row gather
    pub fn tchgeo_tempo_random_walk(row_ptrs: *const i64, num_rows: i64, col_indices: *const i64,
                                    node_timestamps: *const i64, num_node_timestamps: i64, edge_timestamps: *const i64,
                                    start: *const i64, start_timestamps: *const i64, num_walks: i64, walk_length: i64,
                                    window_lo: i64, window_hi: i64, seed: u64, walker_base: i64, walks: *mut i64,
                                    walks_timestamps: *mut i64, scratch: *mut i32, stream: *mut c_void) -> tchgeo_status;
    pub fn tchgeo_negative_sampling_capacity(args: *const tchgeo_negative_args, samples_cap: *mut i64,
                                             edges_cap: *mut i64) -> tchgeo_status;
    pub fn tchgeo_negative_sampling_workspace_bytes(args: *const tchgeo_negative_args) -> usize;
    pub fn tchgeo_negative_sampling(args: *const tchgeo_negative_args) -> tchgeo_status;
    pub fn tchgeo_csc_edge_cumsum_f64(col_ptrs: *const i64, n_cols: i64, row_data: *mut f64, numel: i64,
                                      scratch: *mut i32, stream: *mut c_void) -> tchgeo_status;
    pub fn tchgeo_csc_sort_edges_workspace_bytes(numel: i64, n_cols: i64) -> usize;
    pub fn tchgeo_csc_sort_edges(col_ptrs: *const i64, n_cols: i64, perm: *const i64, row_weights: *const f64,
                                 numel: i64, descending: i32, new_perm: *mut i64, workspace: *mut c_void,
                                 workspace_bytes: usize, stream: *mut c_void) -> tchgeo_status;
    pub fn tchgeo_gather_rows(src: *const c_void, num_rows: i64, row_bytes: i64, index: *const i64, n: i64,
                              dst: *mut c_void, scratch: *mut i32, stream: *mut c_void) -> tchgeo_status;
}

/// nonzero status -> TensorConversionError::Unknown(last_error) -> PyValueError (src/utils/tensor.rs:22-27)
pub fn check(status: tchgeo_status) -> crate::utils::TensorResult<()> {
    if status == TCHGEO_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(tchgeo_last_error()) }.to_string_lossy().into_owned();
    Err(crate::utils::TensorConversionError::Unknown(msg))
}
