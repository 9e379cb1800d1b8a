"""ctypes binding of libtchgeo_cuda.so (the C ABI in include/tchgeo_cuda.h).

This is the Python counterpart of the `extern "C"` block the reference's Rust host (src/python.rs)
would carry: raw device pointers from `Tensor.data_ptr()` are passed straight through.  There is no
fallback: if the CUDA library is missing, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCHGEO_LIB: another build of the same ABI (A/B measurements of two kernel versions on one box)
LIB_PATH = os.environ.get("TCHGEO_LIB") or os.path.join(_HERE, "libtchgeo_cuda.so")

OK, ERR_BAD_ARG, ERR_CUDA, ERR_CAPACITY, ERR_INDEX, ERR_REFERENCE_PANIC, ERR_INTERNAL = range(7)
SAMPLER_UNIFORM, SAMPLER_UNIFORM_REPLACE, SAMPLER_WEIGHTED = 0, 1, 2
ABI_VERSION = 6
PREPARE_INDEX_REPLICA, PREPARE_WEIGHT_RECORDS = 1, 2

c_i64, c_i32, c_u64, c_u32, c_vp, c_sz = (ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint32,
                                          ctypes.c_void_p, ctypes.c_size_t)


class SamplingArgs(ctypes.Structure):
    """struct tchgeo_sampling_args (field order must match include/tchgeo_cuda.h)."""
    _fields_ = [
        ("num_node_types", c_i32), ("num_rels", c_i32), ("num_hops", c_i32), ("sampler_kind", c_i32),
        ("rel_src", c_vp), ("rel_dst", c_vp),
        ("col_ptrs", c_vp), ("num_cols", c_vp), ("row_indices", c_vp), ("weights", c_vp), ("nnz", c_vp),
        ("graph", c_vp),
        ("fanouts", c_vp), ("rel_active", c_vp),
        ("num_batches", c_i64), ("inputs", c_vp), ("seeds_per_batch", c_vp),
        ("seed", c_u64), ("batch_base", c_u32), ("reserved0", c_u32),
        ("samples", c_vp), ("samples_stride", c_vp),
        ("rows", c_vp), ("cols", c_vp), ("edge_index", c_vp), ("edges_stride", c_vp),
        ("samples_len", c_vp), ("edges_len", c_vp), ("layer_offsets", c_vp),
        ("nodes", c_vp), ("local", c_vp), ("nodes_len", c_vp),
        ("filter_mode", c_i32), ("filter_forward", c_i32), ("filter_window_lo", c_i64), ("filter_window_hi", c_i64),
        ("timestamps", c_vp), ("inputs_state", c_vp), ("states", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz), ("stream", c_vp),
    ]


class NegativeArgs(ctypes.Structure):
    """struct tchgeo_negative_args"""
    _fields_ = [
        ("num_node_types", c_i32), ("num_rels", c_i32), ("rel_src", c_vp), ("rel_dst", c_vp),
        ("row_ptrs", c_vp), ("col_indices", c_vp), ("num_rows", c_vp), ("node_count", c_vp),
        ("inputs", c_vp), ("num_inputs", c_vp), ("num_neg", c_i64), ("try_count", c_i64),
        ("inbound", c_i32), ("reserved0", c_i32), ("seed", c_u64),
        ("samples", c_vp), ("rows", c_vp), ("cols", c_vp), ("samples_len", c_vp), ("edges_len", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_sz), ("stream", c_vp),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python tch-geometric_b200/build.py` "
            "(nvcc, sm_100a). tch_geometric (B200) has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.tchgeo_abi_version.restype = c_i32
    lib.tchgeo_last_error.restype = ctypes.c_char_p
    lib.tchgeo_device_set_l2_fetch_granularity.restype = c_i32
    lib.tchgeo_device_set_l2_fetch_granularity.argtypes = [c_i32, c_vp]
    lib.tchgeo_ind2ptr.restype = c_i32
    lib.tchgeo_ind2ptr.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp]
    lib.tchgeo_coo_to_csx_workspace_bytes.restype = c_sz
    lib.tchgeo_coo_to_csx_workspace_bytes.argtypes = [c_i64, c_i64, c_i64]
    lib.tchgeo_coo_to_csx.restype = c_i32
    lib.tchgeo_coo_to_csx.argtypes = [c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_csc_edge_cumsum_f64.restype = c_i32
    lib.tchgeo_csc_edge_cumsum_f64.argtypes = [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]
    lib.tchgeo_csc_sort_edges_workspace_bytes.restype = c_sz
    lib.tchgeo_csc_sort_edges_workspace_bytes.argtypes = [c_i64, c_i64]
    lib.tchgeo_csc_sort_edges.restype = c_i32
    lib.tchgeo_csc_sort_edges.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_graph_create.restype = c_i32
    lib.tchgeo_graph_create.argtypes = [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_graph_set_weights.restype = c_i32
    lib.tchgeo_graph_set_weights.argtypes = [c_vp, c_vp]
    lib.tchgeo_graph_set_timestamps.restype = c_i32
    lib.tchgeo_graph_set_timestamps.argtypes = [c_vp, c_vp]
    lib.tchgeo_graph_prepare.restype = c_i32
    lib.tchgeo_graph_prepare.argtypes = [c_vp, c_i32, c_vp]
    lib.tchgeo_graph_derived_bytes.restype = c_sz
    lib.tchgeo_graph_derived_bytes.argtypes = [c_vp]
    lib.tchgeo_graph_destroy.restype = None
    lib.tchgeo_graph_destroy.argtypes = [c_vp]
    P = ctypes.POINTER(SamplingArgs)
    lib.tchgeo_plan_create.restype = c_i32
    lib.tchgeo_plan_create.argtypes = [P, c_vp]
    lib.tchgeo_plan_enqueue.restype = c_i32
    lib.tchgeo_plan_enqueue.argtypes = [c_vp, c_u64, c_u32, c_vp]
    lib.tchgeo_plan_enqueue_timed.restype = c_i32
    lib.tchgeo_plan_enqueue_timed.argtypes = [c_vp, c_u64, c_u32, c_vp, c_vp, c_i32, c_vp]
    lib.tchgeo_plan_collect.restype = c_i32
    lib.tchgeo_plan_collect.argtypes = [c_vp]
    lib.tchgeo_plan_results.restype = c_i32
    lib.tchgeo_plan_results.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_plan_num_launches.restype = c_i32
    lib.tchgeo_plan_num_launches.argtypes = [c_vp]
    lib.tchgeo_plan_destroy.restype = None
    lib.tchgeo_plan_destroy.argtypes = [c_vp]
    lib.tchgeo_random_walk_graph.restype = c_i32
    lib.tchgeo_random_walk_graph.argtypes = [c_vp, c_i32, c_vp, c_i64, c_i64, ctypes.c_float, ctypes.c_float, c_u64, c_i64,
                                             c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_unique_relabel_batched_workspace_bytes.restype = c_sz
    lib.tchgeo_unique_relabel_batched_workspace_bytes.argtypes = [c_i64, c_i64, c_i64]
    lib.tchgeo_unique_relabel_batched.restype = c_i32
    lib.tchgeo_unique_relabel_batched.argtypes = [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz,
                                                  c_vp, c_vp]
    lib.tchgeo_pack_transport.restype = c_i32
    lib.tchgeo_pack_transport.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp,
                                          c_i64, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_host_unpack_transport.restype = c_i32
    lib.tchgeo_host_unpack_transport.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i32]
    lib.tchgeo_neighbor_sampling_capacity.restype = c_i32
    lib.tchgeo_neighbor_sampling_capacity.argtypes = [P, c_vp, c_vp]
    lib.tchgeo_neighbor_sampling_workspace_bytes.restype = c_sz
    lib.tchgeo_neighbor_sampling_workspace_bytes.argtypes = [P]
    lib.tchgeo_neighbor_sampling.restype = c_i32
    lib.tchgeo_neighbor_sampling.argtypes = [P]
    lib.tchgeo_compress_indices.restype = c_i32
    lib.tchgeo_compress_indices.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp]
    lib.tchgeo_neighbor_sampling_timed.restype = c_i32
    lib.tchgeo_neighbor_sampling_timed.argtypes = [P, c_vp, c_i32, c_vp]
    lib.tchgeo_neighbor_sampling_collect.restype = c_i32
    lib.tchgeo_neighbor_sampling_collect.argtypes = [P]
    lib.tchgeo_neighbor_sampling_homogenous.restype = c_i32
    lib.tchgeo_neighbor_sampling_homogenous.argtypes = [
        c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_i32, c_i32, c_vp, c_u64, c_u32,
        c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_serve_requests.restype = c_i32
    lib.tchgeo_serve_requests.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_i32, c_u64,
                                          c_u32, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_part_begin_hop.restype = c_i32
    lib.tchgeo_part_begin_hop.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_u32, c_vp, c_vp, c_vp, c_vp,
                                          c_vp, c_sz, c_vp]
    lib.tchgeo_serve_requests_rows.restype = c_i32
    lib.tchgeo_serve_requests_rows.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i32, c_u64, c_u32,
                                               c_vp, c_vp, c_vp]
    lib.tchgeo_part_count_hop.restype = c_i32
    lib.tchgeo_part_count_hop.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_part_scatter_hop.restype = c_i32
    lib.tchgeo_part_scatter_hop.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_u32, c_vp, c_vp, c_vp, c_vp,
                                            c_vp, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_serve_requests_rows_peer.restype = c_i32
    lib.tchgeo_serve_requests_rows_peer.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i32, c_u64,
                                                    c_u32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_part_hop_workspace_bytes.restype = c_sz
    lib.tchgeo_part_hop_workspace_bytes.argtypes = [c_i64, c_i64]
    lib.tchgeo_part_finish_hop.restype = c_i32
    lib.tchgeo_part_finish_hop.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp,
                                           c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_partf_workspace_bytes.restype = c_sz
    lib.tchgeo_partf_workspace_bytes.argtypes = [c_i64, c_i64]
    lib.tchgeo_partf_scatter.restype = c_i32
    lib.tchgeo_partf_scatter.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_u32, c_i64, c_vp, c_vp,
                                         c_vp, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_partf_serve.restype = c_i32
    lib.tchgeo_partf_serve.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32,
                                       c_u64, c_u32, c_i32, c_i32, c_vp, c_vp, c_vp]
    lib.tchgeo_partf_finish.restype = c_i32
    lib.tchgeo_partf_finish.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]
    lib.tchgeo_status_from_error_word.restype = c_i32
    lib.tchgeo_status_from_error_word.argtypes = [c_u32]
    lib.tchgeo_random_walk.restype = c_i32
    lib.tchgeo_random_walk.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, ctypes.c_float, ctypes.c_float, c_u64,
                                       c_i64, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_random_walk_ex.restype = c_i32
    lib.tchgeo_random_walk_ex.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, ctypes.c_float, ctypes.c_float,
                                          c_u64, c_i64, c_vp, c_vp, c_vp, c_vp]
    PN = ctypes.POINTER(NegativeArgs)
    lib.tchgeo_negative_sampling_capacity.restype = c_i32
    lib.tchgeo_negative_sampling_capacity.argtypes = [PN, c_vp, c_vp]
    lib.tchgeo_negative_sampling_workspace_bytes.restype = c_sz
    lib.tchgeo_negative_sampling_workspace_bytes.argtypes = [PN]
    lib.tchgeo_negative_sampling.restype = c_i32
    lib.tchgeo_negative_sampling.argtypes = [PN]
    lib.tchgeo_tempo_random_walk.restype = c_i32
    lib.tchgeo_tempo_random_walk.argtypes = [c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_u64,
                                             c_i64, c_vp, c_vp, c_vp, c_vp]
    lib.tchgeo_gather_rows.restype = c_i32
    lib.tchgeo_gather_rows.argtypes = [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]
    lib.tchgeo_pack_ragged.restype = c_i32
    lib.tchgeo_pack_ragged.argtypes = [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]
    lib.tchgeo_unique_relabel_workspace_bytes.restype = c_sz
    lib.tchgeo_unique_relabel_workspace_bytes.argtypes = [c_i64]
    lib.tchgeo_unique_relabel.restype = c_i32
    lib.tchgeo_unique_relabel.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]
    if lib.tchgeo_abi_version() != ABI_VERSION:
        raise ImportError("libtchgeo_cuda.so ABI version mismatch")
    return lib


lib = _load()

EXPORTS = [
    "tchgeo_abi_version", "tchgeo_last_error", "tchgeo_device_set_l2_fetch_granularity", "tchgeo_ind2ptr", "tchgeo_coo_to_csx_workspace_bytes",
    "tchgeo_coo_to_csx", "tchgeo_csc_edge_cumsum_f64", "tchgeo_csc_sort_edges_workspace_bytes", "tchgeo_csc_sort_edges", "tchgeo_compress_indices", "tchgeo_neighbor_sampling_capacity", "tchgeo_neighbor_sampling_workspace_bytes",
    "tchgeo_neighbor_sampling", "tchgeo_neighbor_sampling_timed", "tchgeo_neighbor_sampling_collect", "tchgeo_neighbor_sampling_homogenous",
    "tchgeo_serve_requests", "tchgeo_part_begin_hop", "tchgeo_part_count_hop", "tchgeo_part_scatter_hop", "tchgeo_serve_requests_rows", "tchgeo_serve_requests_rows_peer", "tchgeo_part_hop_workspace_bytes", "tchgeo_part_finish_hop", "tchgeo_status_from_error_word", "tchgeo_negative_sampling_capacity", "tchgeo_negative_sampling_workspace_bytes", "tchgeo_negative_sampling", "tchgeo_random_walk", "tchgeo_random_walk_ex", "tchgeo_tempo_random_walk", "tchgeo_gather_rows", "tchgeo_unique_relabel_workspace_bytes", "tchgeo_unique_relabel",
    "tchgeo_unique_relabel_batched_workspace_bytes", "tchgeo_unique_relabel_batched",
    "tchgeo_graph_create", "tchgeo_graph_set_weights", "tchgeo_graph_set_timestamps", "tchgeo_graph_prepare",
    "tchgeo_graph_derived_bytes", "tchgeo_graph_destroy",
    "tchgeo_plan_create", "tchgeo_plan_enqueue", "tchgeo_plan_enqueue_timed", "tchgeo_plan_collect", "tchgeo_plan_results",
    "tchgeo_plan_num_launches", "tchgeo_plan_destroy", "tchgeo_random_walk_graph", "tchgeo_pack_ragged",
    "tchgeo_partf_workspace_bytes", "tchgeo_partf_scatter", "tchgeo_partf_serve", "tchgeo_partf_finish",
    "tchgeo_pack_transport", "tchgeo_host_unpack_transport",
]


def last_error():
    return lib.tchgeo_last_error().decode("utf-8", "replace")


class ReferencePanic(RuntimeError):
    """Input on which the reference panics (pyo3 PanicException there)."""


def check(status):
    """Map a tchgeo_status to the reference's error behaviour: TensorConversionError -> ValueError
    (src/utils/tensor.rs:22-27); inputs the reference panics on -> ReferencePanic."""
    if status == OK:
        return
    msg = last_error()
    if status in (ERR_INDEX, ERR_REFERENCE_PANIC):
        raise ReferencePanic(msg)
    if status == ERR_CAPACITY:
        raise MemoryError(msg)
    if status in (ERR_CUDA, ERR_INTERNAL):
        raise RuntimeError(msg)
    raise ValueError(msg)
