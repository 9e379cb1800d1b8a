//! `extern "C"` declarations of include/tchgeo_cuda.h (ABI version 6) for the reference crate: new file
//! `src/cuda_ffi.rs`.  One declaration per entry point of the header, same order.
//! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo/rustc); tests/test_abi.py checks the same layout through ctypes
//! and tests/cpp/abi_harness.cpp drives the same calls from C++.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

pub type tchgeo_status = i32;
pub const TCHGEO_ABI_VERSION: i32 = 6;
pub const TCHGEO_OK: tchgeo_status = 0;
pub const TCHGEO_ERR_BAD_ARG: tchgeo_status = 1;
pub const TCHGEO_ERR_CUDA: tchgeo_status = 2;
pub const TCHGEO_ERR_CAPACITY: tchgeo_status = 3;
pub const TCHGEO_ERR_INDEX: tchgeo_status = 4;
pub const TCHGEO_ERR_REFERENCE_PANIC: tchgeo_status = 5;
pub const TCHGEO_ERR_INTERNAL: tchgeo_status = 6;
pub const TCHGEO_SAMPLER_UNIFORM: i32 = 0;
pub const TCHGEO_SAMPLER_UNIFORM_REPLACE: i32 = 1;
pub const TCHGEO_SAMPLER_WEIGHTED: i32 = 2;
pub const TCHGEO_PREPARE_INDEX_REPLICA: i32 = 1;
pub const TCHGEO_PREPARE_WEIGHT_RECORDS: i32 = 2;

pub type tchgeo_stream = *mut c_void;
#[repr(C)] pub struct tchgeo_graph_t { _private: [u8; 0] }
#[repr(C)] pub struct tchgeo_plan_t { _private: [u8; 0] }

#[repr(C)]
pub struct tchgeo_sampling_args {
    pub num_node_types: i32, pub num_rels: i32, pub num_hops: i32, pub sampler_kind: i32,
    pub rel_src: *const i32, pub rel_dst: *const i32,
    pub col_ptrs: *const *const i64, pub num_cols: *const i64, pub row_indices: *const *const i64,
    pub weights: *const *const f64, pub nnz: *const i64, pub graph: *const tchgeo_graph_t,
    pub fanouts: *const i64, pub rel_active: *const u8,
    pub num_batches: i64, pub inputs: *const *const i64, pub seeds_per_batch: *const i64,
    pub seed: u64, pub batch_base: u32, pub reserved0: u32,
    pub samples: *const *mut i64, pub samples_stride: *const i64,
    pub rows: *const *mut i64, pub cols: *const *mut i64, pub edge_index: *const *mut i64,
    pub edges_stride: *const i64,
    pub samples_len: *mut i64, pub edges_len: *mut i64, pub layer_offsets: *mut i64,
    pub nodes: *const *mut i64, pub local: *const *mut i64, pub nodes_len: *mut i64,
    pub filter_mode: i32, pub filter_forward: i32, pub filter_window_lo: i64, pub filter_window_hi: i64,
    pub timestamps: *const *const i64, pub inputs_state: *const *const i64, pub states: *const *mut i64,
    pub workspace: *mut c_void, pub workspace_bytes: usize, pub stream: tchgeo_stream,
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
    pub workspace: *mut c_void, pub workspace_bytes: usize, pub stream: tchgeo_stream,
}

extern "C" {
    pub fn tchgeo_abi_version() -> i32;
    pub fn tchgeo_last_error() -> *const c_char;
    pub fn tchgeo_device_set_l2_fetch_granularity(bytes: i32, actual: *mut i32) -> tchgeo_status;
    // src/data/storage.rs:67-127
    pub fn tchgeo_ind2ptr(ind: *const i64, numel: i64, m: i64, out: *mut i64, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_coo_to_csx_workspace_bytes(num_edges: i64, n_rows: i64, n_cols: i64) -> usize;
    pub fn tchgeo_coo_to_csx(row: *const i64, col: *const i64, num_edges: i64, n_rows: i64, n_cols: i64, csc: i32,
                             ptrs: *mut i64, indices: *mut i64, perm: *mut i64, workspace: *mut c_void,
                             workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    // src/data/transform.rs
    pub fn tchgeo_csc_edge_cumsum_f64(col_ptrs: *const i64, n_cols: i64, row_data: *mut f64, numel: i64,
                                      scratch: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_csc_sort_edges_workspace_bytes(numel: i64, n_cols: i64) -> usize;
    pub fn tchgeo_csc_sort_edges(col_ptrs: *const i64, n_cols: i64, perm: *const i64, row_weights: *const f64, numel: i64,
                                 descending: i32, new_perm: *mut i64, workspace: *mut c_void, workspace_bytes: usize,
                                 stream: tchgeo_stream) -> tchgeo_status;
    // graph handle (borrowed CSC/CSR arrays + derived int32 replica / weight records)
    pub fn tchgeo_graph_create(num_rels: i32, ptrs: *const *const i64, num_major: *const i64, indices: *const *const i64,
                               nnz: *const i64, out: *mut *mut tchgeo_graph_t) -> tchgeo_status;
    pub fn tchgeo_graph_set_weights(g: *mut tchgeo_graph_t, weights: *const *const f64) -> tchgeo_status;
    pub fn tchgeo_graph_set_timestamps(g: *mut tchgeo_graph_t, timestamps: *const *const i64) -> tchgeo_status;
    pub fn tchgeo_graph_prepare(g: *mut tchgeo_graph_t, what: i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_graph_derived_bytes(g: *const tchgeo_graph_t) -> usize;
    pub fn tchgeo_graph_destroy(g: *mut tchgeo_graph_t);
    // src/algo/neighbor_sampling.rs:162-356
    pub fn tchgeo_compress_indices(src: *const i64, n: i64, dst: *mut i32, scratch: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_capacity(args: *const tchgeo_sampling_args, samples_cap: *mut i64, edges_cap: *mut i64) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_workspace_bytes(args: *const tchgeo_sampling_args) -> usize;
    pub fn tchgeo_neighbor_sampling(args: *const tchgeo_sampling_args) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_timed(args: *const tchgeo_sampling_args, launch_ms: *mut f32, cap: i32, n: *mut i32) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_collect(args: *const tchgeo_sampling_args) -> tchgeo_status;
    pub fn tchgeo_neighbor_sampling_homogenous(col_ptrs: *const i64, num_cols: i64, row_indices: *const i64, inputs: *const i64,
        num_batches: i64, seeds_per_batch: i64, num_neighbors: *const i64, num_hops: i32, sampler_kind: i32,
        weights: *const f64, seed: u64, batch_base: u32, samples: *mut i64, samples_stride: i64, rows: *mut i64,
        cols: *mut i64, edge_index: *mut i64, edges_stride: i64, out_lens: *mut i64, layer_offsets: *mut i64,
        workspace: *mut c_void, workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    // plan handle
    pub fn tchgeo_plan_create(args: *const tchgeo_sampling_args, out: *mut *mut tchgeo_plan_t) -> tchgeo_status;
    pub fn tchgeo_plan_enqueue(plan: *mut tchgeo_plan_t, seed: u64, batch_base: u32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_plan_enqueue_timed(plan: *mut tchgeo_plan_t, seed: u64, batch_base: u32, stream: tchgeo_stream,
                                     launch_ms: *mut f32, cap: i32, n: *mut i32) -> tchgeo_status;
    pub fn tchgeo_plan_collect(plan: *mut tchgeo_plan_t) -> tchgeo_status;
    pub fn tchgeo_plan_results(plan: *const tchgeo_plan_t, samples_len: *mut *const i64, edges_len: *mut *const i64,
                               layer_offsets: *mut *const i64, nodes_len: *mut *const i64) -> tchgeo_status;
    pub fn tchgeo_plan_num_launches(plan: *const tchgeo_plan_t) -> i32;
    pub fn tchgeo_plan_destroy(plan: *mut tchgeo_plan_t);
    // range-partitioned CSC (config 5): owner side, legacy requester pipeline, device-only fixed-segment protocol
    pub fn tchgeo_serve_requests(ptrs_local: *const i64, indices_local: *const i64, weights_local: *const f64, col_begin: i64,
        ncols_local: i64, edge_base: i64, req_ids: *const i64, req_meta: *const i64, n: i64, fanout: i64, sampler_kind: i32,
        seed: u64, rel: u32, out_ids: *mut i64, out_ptrs: *mut i64, err_scratch: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_part_begin_hop(samples: *const i64, samples_stride: i64, fr_begin: *const i64, fr_end: *const i64,
        num_batches: i64, frontier_cap: i64, cols_per_rank: i64, world: i32, batch_base: u32, counts: *mut i64,
        cursor: *mut i64, req: *mut i64, err_word: *mut i32, workspace: *mut c_void, workspace_bytes: usize,
        stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_part_count_hop(samples: *const i64, samples_stride: i64, fr_begin: *const i64, fr_end: *const i64,
        num_batches: i64, frontier_cap: i64, cols_per_rank: i64, world: i32, counts: *mut i64, cursor: *mut i64,
        err_word: *mut i32, workspace: *mut c_void, workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_part_scatter_hop(samples: *const i64, samples_stride: i64, fr_begin: *const i64, fr_end: *const i64,
        num_batches: i64, frontier_cap: i64, cols_per_rank: i64, world: i32, batch_base: u32, counts: *const i64,
        cursor: *mut i64, req: *mut i64, peer_req: *const *mut c_void, peer_row0: *const i64, err_word: *mut i32,
        workspace: *mut c_void, workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_serve_requests_rows(ptrs_local: *const i64, indices_local: *const i64, weights_local: *const f64,
        col_begin: i64, ncols_local: i64, nnz_local: i64, req: *const i64, n: i64, fanout: i64, sampler_kind: i32, seed: u64,
        rel: u32, ans: *mut i32, err_word: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_serve_requests_rows_peer(ptrs_local: *const i64, indices_local: *const i64, weights_local: *const f64,
        col_begin: i64, ncols_local: i64, nnz_local: i64, req: *const i64, n: i64, fanout: i64, sampler_kind: i32, seed: u64,
        rel: u32, world: i32, peer_ans: *const *mut c_void, recv_counts: *const i64, peer_row0: *const i64,
        err_word: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_part_hop_workspace_bytes(num_batches: i64, frontier_cap: i64) -> usize;
    pub fn tchgeo_part_finish_hop(req: *const i64, ans: *const i32, num_requests: i64, fanout: i64, owner_edge_base: *const i64,
        cols_per_rank: i64, world: i32, fr_begin: *const i64, fr_end: *const i64, num_batches: i64, frontier_cap: i64,
        node_len_in: *const i64, edge_len_in: *const i64, node_len_out: *mut i64, edge_len_out: *mut i64, samples: *mut i64,
        samples_stride: i64, rows: *mut i64, cols: *mut i64, edge_index: *mut i64, edges_stride: i64, err_word: *mut i32,
        workspace: *mut c_void, workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_partf_workspace_bytes(num_batches: i64, frontier_cap: i64) -> usize;
    pub fn tchgeo_partf_scatter(samples: *const i64, samples_stride: i64, fr_begin: *const i64, fr_end: *const i64,
        num_batches: i64, frontier_cap: i64, cols_per_rank: i64, world: i32, me: i32, batch_base: u32, seg_rows: i64,
        send: *mut i64, cursor: *mut i64, slot_of: *mut u32, peer_req: *const *mut c_void, peer_cnt: *const *mut c_void,
        err_word: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_partf_serve(ptrs_local: *const i64, indices_local: *const i64, indices32_local: *const i32,
        weights_local: *const f64, col_begin: i64, ncols_local: i64, nnz_local: i64, req_in: *const i64, cnt_in: *const i64,
        seg_rows: i64, max_requests: i64, fanout: i64, sampler_kind: i32, seed: u64, rel: u32, world: i32, me: i32,
        peer_ans: *const *mut c_void, err_word: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_partf_finish(ans_in: *const i32, slot_of: *const u32, seg_rows: i64, fanout: i64, owner_edge_base: *const i64,
        world: i32, fr_begin: *const i64, fr_end: *const i64, num_batches: i64, frontier_cap: i64, node_len_in: *const i64,
        edge_len_in: *const i64, node_len_out: *mut i64, edge_len_out: *mut i64, samples: *mut i64, samples_stride: i64,
        rows: *mut i64, cols: *mut i64, edge_index: *mut i64, edges_stride: i64, err_word: *mut i32, workspace: *mut c_void,
        workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_status_from_error_word(word: u32) -> tchgeo_status;
    // src/algo/random_walk.rs
    pub fn tchgeo_random_walk(row_ptrs: *const i64, num_rows: i64, col_indices: *const i64, start: *const i64, num_walks: i64,
        walk_length: i64, p: f32, q: f32, seed: u64, walker_base: i64, walks: *mut i64, stats: *mut i64,
        attempts_out: *mut i64, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_random_walk_ex(row_ptrs: *const i64, num_rows: i64, col_indices: *const i64, col_indices32: *const i32,
        start: *const i64, num_walks: i64, walk_length: i64, p: f32, q: f32, seed: u64, walker_base: i64, walks: *mut i64,
        stats: *mut i64, attempts_out: *mut i64, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_random_walk_graph(graph: *const tchgeo_graph_t, rel: i32, start: *const i64, num_walks: i64, walk_length: i64,
        p: f32, q: f32, seed: u64, walker_base: i64, walks: *mut i64, stats: *mut i64, attempts_out: *mut i64,
        stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_tempo_random_walk(row_ptrs: *const i64, num_rows: i64, col_indices: *const i64, node_timestamps: *const i64,
        num_node_timestamps: i64, edge_timestamps: *const i64, start: *const i64, start_timestamps: *const i64,
        num_walks: i64, walk_length: i64, window_lo: i64, window_hi: i64, seed: u64, walker_base: i64, walks: *mut i64,
        walks_timestamps: *mut i64, scratch: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    // src/algo/negative_sampling.rs
    pub fn tchgeo_negative_sampling_capacity(args: *const tchgeo_negative_args, samples_cap: *mut i64, edges_cap: *mut i64) -> tchgeo_status;
    pub fn tchgeo_negative_sampling_workspace_bytes(args: *const tchgeo_negative_args) -> usize;
    pub fn tchgeo_negative_sampling(args: *const tchgeo_negative_args) -> tchgeo_status;
    // downstream gather / host transfer / dedup + relabel
    pub fn tchgeo_gather_rows(src: *const c_void, num_rows: i64, row_bytes: i64, index: *const i64, n: i64, dst: *mut c_void,
                              scratch: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_pack_ragged(src: *const i64, stride: i64, lens: *const i64, lens_stride: i64, num_batches: i64, max_len: i64,
                              dst: *mut i64, offsets: *mut i64, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_pack_transport(samples: *const i64, samples_stride: i64, cols: *const i64, edge_index: *const i64,
        edges_stride: i64, n_lens: *const i64, e_lens: *const i64, count: i64, max_n: i64, max_e: i64, samples32: *mut i32,
        eidx32: *mut i32, counts: *mut u8, counts_bytes: i64, n_off: *mut i64, e_off: *mut i64, err_word: *mut i32,
        stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_host_unpack_transport(samples32: *const i32, eidx32: *const i32, counts: *const u8, n_off: *const i64,
        e_off: *const i64, count: i64, samples: *mut i64, cols: *mut i64, edge_index: *mut i64, num_threads: i32) -> tchgeo_status;
    pub fn tchgeo_unique_relabel_batched_workspace_bytes(num_batches: i64, n_max: i64, id_bound: i64) -> usize;
    pub fn tchgeo_unique_relabel_batched(samples: *const i64, stride: i64, lens: *const i64, num_batches: i64, num_seeds: i64,
        n_max: i64, id_bound: i64, nodes: *mut i64, local: *mut i64, nodes_len: *mut i64, workspace: *mut c_void,
        workspace_bytes: usize, err_word: *mut i32, stream: tchgeo_stream) -> tchgeo_status;
    pub fn tchgeo_unique_relabel_workspace_bytes(n: i64) -> usize;
    pub fn tchgeo_unique_relabel(samples: *const i64, n: i64, num_seeds: i64, nodes: *mut i64, local: *mut i64,
                                 num_nodes: *mut i64, workspace: *mut c_void, workspace_bytes: usize, stream: tchgeo_stream) -> tchgeo_status;
    /// csrc/torch_stream_shim.cpp: the raw cudaStream_t of torch's CURRENT stream on `device_index`
    pub fn tchgeo_torch_current_stream(device_index: i32) -> tchgeo_stream;
}

/// tchgeo_status -> the reference's error type: TensorConversionError -> PyValueError (src/utils/tensor.rs:10-27);
/// inputs on which the reference panics keep panicking (pyo3 turns that into PanicException).
pub fn check(st: tchgeo_status) -> Result<(), crate::utils::TensorConversionError> {
    if st == TCHGEO_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(tchgeo_last_error()) }.to_string_lossy().into_owned();
    if st == TCHGEO_ERR_INDEX || st == TCHGEO_ERR_REFERENCE_PANIC { panic!("{}", msg); }
    Err(crate::utils::TensorConversionError::Unknown(msg))
}
