/*
 * tchgeo_cuda.h -- C ABI of the B200-native (sm_100a) sampling hot path for tch-geometric.
 *
 * This is the drop-in boundary: the entry points below are what the reference's Rust host
 * (src/python.rs) binds through `extern "C"` in place of its CPU algorithms.  Every pointer marked
 * DEVICE is a CUDA device pointer taken straight from `tch::Tensor::data_ptr()` (or
 * `torch.Tensor.data_ptr()`); every pointer marked HOST is ordinary host memory.  No torch types,
 * no C++ types, no exceptions: every entry returns a tchgeo_status (0 = ok).  The library never
 * allocates user-visible memory: outputs and workspaces are allocated by the caller (sizes come
 * from the *_workspace_bytes / *_capacity queries) and narrowed to the returned lengths.
 * There is no CPU fallback behind this ABI.
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   tchgeo_ind2ptr                      src/data/storage.rs:67-101   (ind2ptr)
 *   tchgeo_coo_to_csx                   src/data/storage.rs:103-127  (TryFrom<&CooGraphStorage>),
 *                                       called from src/python.rs:27-39 (to_csc) and :41-53 (to_csr)
 *   tchgeo_neighbor_sampling_*          src/algo/neighbor_sampling.rs:162-230 (homogenous) and
 *                                       :233-356 (heterogenous), called from src/python.rs:251,363;
 *                                       samplers src/utils/sampling.rs:6-69
 *   tchgeo_random_walk                  src/algo/random_walk.rs:10-75, called from src/python.rs:598
 *   tchgeo_unique_relabel               insertion-order relabel map, semantic of
 *                                       src/algo/negative_sampling.rs:20-47 (the dedup stage
 *                                       north_star asks for; additive, never alters drop-in outputs)
 *
 * Threading: all entry points are re-entrant and keep no global mutable state except the
 * thread-local last-error string.  Work is issued on the caller's stream.
 *
 * RNG contract: Philox4x32-10, key = (seed_lo, seed_hi); counters are documented in DESIGN.md
 * ("RNG contract") and are independent of launch geometry, so results depend only on
 * (seed, batch index, relation index, position of the frontier node in its samples vector).
 */
#ifndef TCHGEO_CUDA_H_
#define TCHGEO_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCHGEO_ABI_VERSION 6

#if defined(__GNUC__)
#define TCHGEO_API __attribute__((visibility("default")))
#else
#define TCHGEO_API
#endif

typedef int32_t tchgeo_status;
enum {
  TCHGEO_OK = 0,
  TCHGEO_ERR_BAD_ARG = 1,       /* null pointer, negative size, unsupported fanout ...            */
  TCHGEO_ERR_CUDA = 2,          /* a CUDA runtime call failed; see tchgeo_last_error()            */
  TCHGEO_ERR_CAPACITY = 3,      /* an output buffer or the workspace is too small                 */
  TCHGEO_ERR_INDEX = 4,         /* node id / edge endpoint out of range (reference: Rust panic)   */
  TCHGEO_ERR_REFERENCE_PANIC = 5, /* input on which the reference panics (fanout 0 with deg > 0,
                                     non-positive weight sum, src/utils/sampling.rs:19,49,51)      */
  TCHGEO_ERR_INTERNAL = 6       /* look-back watchdog tripped (should never happen)               */
};

/* sampler kinds: src/python.rs:210-216 dispatch */
enum {
  TCHGEO_SAMPLER_UNIFORM = 0,         /* UnweightedSampler<false> (default), reservoir, quirk Q1 */
  TCHGEO_SAMPLER_UNIFORM_REPLACE = 1, /* UnweightedSampler<true>, exactly k iid picks, quirk Q3   */
  TCHGEO_SAMPLER_WEIGHTED = 2         /* WeightedSampler<f64>, quirk Q2                           */
};

/* cudaStream_t passed as void* so that C callers need no CUDA headers. */
typedef void* tchgeo_stream;

TCHGEO_API int32_t tchgeo_abi_version(void);
/* Human-readable description of the last non-OK status on this thread (never NULL). */
TCHGEO_API const char* tchgeo_last_error(void);

/* Device tuning hint (optional, process-wide for the current device): L2 -> DRAM fetch granularity in
 * bytes (32, 64 or 128; cudaLimitMaxL2FetchGranularity).  The sampling path reads one 32-byte sector per
 * random 8-byte gather, so 32 avoids fetching neighbour sectors that are never used.  `actual` (HOST,
 * may be NULL) receives the value the driver reports afterwards. */
TCHGEO_API tchgeo_status tchgeo_device_set_l2_fetch_granularity(int32_t bytes, int32_t* actual);

/* -------------------------------------------------------------------------------------------- */
/* ind2ptr: out[i] = #{e : ind[e] < i}, i in [0, m]; `ind` sorted ascending.                      */
/* replaces src/data/storage.rs:67-101                                                           */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API tchgeo_status tchgeo_ind2ptr(const int64_t* ind /*DEVICE [numel]*/, int64_t numel, int64_t m,
                             int64_t* out /*DEVICE [m+1]*/, tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* COO -> CSC (csc != 0) or CSR (csc == 0).  perm = argsort(major*size_minor + minor) (stable),   */
/* ptrs = ind2ptr(major[perm]), indices = minor[perm].  n_rows = size.0, n_cols = size.1.          */
/* replaces src/data/storage.rs:103-127                                                          */
/* Two forms give the same three arrays: buckets of consecutive major ids finished in shared       */
/* memory (the default where the buckets fit), and a device-wide radix sort of (major, minor) keys  */
/* (everything else; the environment variable TCHGEO_CSX_SORT=cub | partition forces one).  The     */
/* call synchronises `stream` before it returns (status of the ids it read).                        */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API size_t tchgeo_coo_to_csx_workspace_bytes(int64_t num_edges, int64_t n_rows, int64_t n_cols);
TCHGEO_API tchgeo_status tchgeo_coo_to_csx(const int64_t* row /*DEVICE [E]*/, const int64_t* col /*DEVICE [E]*/,
                                int64_t num_edges, int64_t n_rows, int64_t n_cols, int32_t csc,
                                int64_t* ptrs /*DEVICE [n_major+1]*/, int64_t* indices /*DEVICE [E]*/,
                                int64_t* perm /*DEVICE [E]*/, void* workspace /*DEVICE*/, size_t workspace_bytes,
                                tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* Per-column transforms of CSC edge data.                                       */
/* tchgeo_csc_edge_cumsum_f64: in-place inclusive prefix sum inside every column, summed serially in  */
/* CSC order; columns with <= 1 element are skipped and a column end beyond numel is clamped, as      */
/* Tensor::slice does.  replaces src/data/transform.rs:36-60 (f64).                                  */
/* tchgeo_csc_sort_edges: new_perm = perm with every column's entries reordered by ascending (or      */
/* descending) weight; stable (ties keep CSC order; the reference's torch argsort is unstable).       */
/* replaces src/data/transform.rs:7-34.                                                              */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API tchgeo_status tchgeo_csc_edge_cumsum_f64(const int64_t* col_ptrs /*DEVICE [n_cols+1]*/, int64_t n_cols,
                                                    double* row_data /*DEVICE [numel], in place*/, int64_t numel,
                                                    int32_t* scratch /*DEVICE [1]*/, tchgeo_stream stream);
TCHGEO_API size_t tchgeo_csc_sort_edges_workspace_bytes(int64_t numel, int64_t n_cols);
TCHGEO_API tchgeo_status tchgeo_csc_sort_edges(const int64_t* col_ptrs /*DEVICE [n_cols+1]*/, int64_t n_cols,
                                               const int64_t* perm /*DEVICE [numel]*/,
                                               const double* row_weights /*DEVICE [numel]*/, int64_t numel,
                                               int32_t descending, int64_t* new_perm /*DEVICE [numel]*/,
                                               void* workspace /*DEVICE*/, size_t workspace_bytes, tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* Graph handle: R relations' CSC (or CSR) arrays plus the DERIVED device arrays the fast paths read -- the      */
/* int32 replica of the indices (random gathers then span half as many DRAM lines) and the interleaved         */
/* (weight, per-column prefix sum) records of the weighted sampler (one 16-byte load per step; the prefix sums   */
/* are src/data/transform.rs:36-60 summed serially, i.e. the reference's own w_sum sequence,                     */
/* src/utils/sampling.rs:37-48).  The caller's arrays are BORROWED (they must outlive the handle and must not    */
/* change); derived arrays are owned by the handle (cudaMalloc, freed by tchgeo_graph_destroy) and are built by   */
/* tchgeo_graph_prepare or on first use.  A handle is what a host keeps per graph in place of the reference's     */
/* borrowed CscGraph views (src/data/graph.rs:33-50, built per call at src/python.rs:204-206, :294-300).          */
/* Not thread-safe while it is being prepared; read-only afterwards.                                             */
/* -------------------------------------------------------------------------------------------- */
typedef struct tchgeo_graph tchgeo_graph_t;
enum { TCHGEO_PREPARE_INDEX_REPLICA = 1, TCHGEO_PREPARE_WEIGHT_RECORDS = 2 };
TCHGEO_API tchgeo_status tchgeo_graph_create(int32_t num_rels, const int64_t* const* ptrs /*HOST [R] of DEVICE [num_major[r]+1]*/,
                                             const int64_t* num_major /*HOST [R] columns (CSC) or rows (CSR)*/,
                                             const int64_t* const* indices /*HOST [R] of DEVICE [nnz[r]]*/,
                                             const int64_t* nnz /*HOST [R]*/, tchgeo_graph_t** out);
/* Per-entry edge data aligned with `indices` (entries may be NULL): sampler weights (f64) / edge timestamps (i64).
 * Setting new weights drops the derived weight records. */
TCHGEO_API tchgeo_status tchgeo_graph_set_weights(tchgeo_graph_t* g, const double* const* weights /*HOST [R] of DEVICE [nnz[r]]*/);
TCHGEO_API tchgeo_status tchgeo_graph_set_timestamps(tchgeo_graph_t* g, const int64_t* const* timestamps /*HOST [R]*/);
/* Build the derived arrays named by `what` (TCHGEO_PREPARE_* bits) now; synchronises the stream.  Relations whose ids
 * do not fit int32 simply keep no replica (not an error). */
TCHGEO_API tchgeo_status tchgeo_graph_prepare(tchgeo_graph_t* g, int32_t what, tchgeo_stream stream);
TCHGEO_API size_t tchgeo_graph_derived_bytes(const tchgeo_graph_t* g);
TCHGEO_API void tchgeo_graph_destroy(tchgeo_graph_t* g);

/* -------------------------------------------------------------------------------------------- */
/* Multi-hop neighbor sampling.  One call samples `num_batches` independent seed batches (the     */
/* reference call is num_batches == 1); batch b's outputs are exactly what the reference returns   */
/* for inputs[b] and live at offset b*stride of each output buffer.                                */
/* -------------------------------------------------------------------------------------------- */
typedef struct tchgeo_sampling_args {
  /* ---- graph: node types, relations (canonical relation order = array order, quirk Q6) ---- */
  int32_t num_node_types;           /* T (1 for homogeneous)                                      */
  int32_t num_rels;                 /* R (1 for homogeneous)                                      */
  int32_t num_hops;                 /* H                                                          */
  int32_t sampler_kind;             /* TCHGEO_SAMPLER_*                                           */
  const int32_t* rel_src;           /* HOST [R] node-type index of the relation's source          */
  const int32_t* rel_dst;           /* HOST [R] node-type index of the relation's destination     */
  const int64_t* const* col_ptrs;   /* HOST [R] of DEVICE [num_cols[r]+1]                          */
  const int64_t* num_cols;          /* HOST [R] number of dst nodes of the relation                */
  const int64_t* const* row_indices;/* HOST [R] of DEVICE [nnz_r]                                  */
  const double* const* weights;     /* HOST [R] of DEVICE [nnz_r] (f64) or NULL unless WEIGHTED    */
  const int64_t* nnz;              /* HOST [R] entries of row_indices[r] (and of weights[r] / timestamps[r]); a column
                                       that ends beyond it raises TCHGEO_ERR_INDEX (the reference panics on the slice,
                                       src/data/graph.rs:72-78).  NULL = unchecked.                                */
  const tchgeo_graph_t* graph;      /* optional: take col_ptrs / num_cols / row_indices / nnz / weights / timestamps from
                                       this handle (the tables above and `timestamps` below may then be NULL) and use its
                                       derived arrays: int32 gathers, and weighted sampling on (weight, prefix sum)
                                       records, which is bit-exact for arbitrary weights.  Outputs do not change.  */
  const int64_t* fanouts;           /* HOST [R*H] num_neighbors[r][hop]                            */
  const uint8_t* rel_active;        /* HOST [R] 0 = relation absent from num_neighbors, or NULL    */
  /* ---- seeds ---- */
  int64_t num_batches;              /* B                                                          */
  const int64_t* const* inputs;     /* HOST [T] of DEVICE [B, seeds_per_batch[t]] (may be NULL if 0) */
  const int64_t* seeds_per_batch;   /* HOST [T]                                                    */
  /* ---- rng ---- */
  uint64_t seed;
  uint32_t batch_base;              /* batch b draws with batch index batch_base + b               */
  uint32_t reserved0;
  /* ---- outputs (caller allocated) ---- */
  int64_t* const* samples;          /* HOST [T] of DEVICE [B, samples_stride[t]]                   */
  const int64_t* samples_stride;    /* HOST [T] per-batch capacity                                 */
  int64_t* const* rows;             /* HOST [R] of DEVICE [B, edges_stride[r]]                     */
  int64_t* const* cols;             /* HOST [R]                                                    */
  int64_t* const* edge_index;       /* HOST [R]                                                    */
  const int64_t* edges_stride;      /* HOST [R] per-batch capacity                                 */
  /* ---- host results, filled after one stream synchronisation (all may be NULL to skip) ---- */
  int64_t* samples_len;             /* HOST [B*T]                                                  */
  int64_t* edges_len;               /* HOST [B*R]                                                  */
  int64_t* layer_offsets;           /* HOST [B*R*H*3] LayerOffset = (len(samples[src]),
                                       len(edges[rel]), len(samples[dst])) at relation start,
                                       src/algo/neighbor_sampling.rs:193,:314-315                  */
  /* ---- dedup + insertion-order relabel of every batch's tree (K7; semantic of src/algo/negative_sampling.rs:20-47),
   *      an ADDITIVE stage enqueued after the hops: never alters the outputs above.  All NULL = skipped. ---- */
  int64_t* const* nodes;            /* HOST [T] of DEVICE [B, samples_stride[t]]: seeds (duplicates kept) ++ every other id
                                       of samples[t] at its first appearance                                        */
  int64_t* const* local;            /* HOST [T] of DEVICE [B, samples_stride[t]]: local[i] = index into nodes of
                                       samples[t][i] (a duplicated seed maps to its LAST seed slot); relabelled edges of
                                       relation r are (local[src][rows[e]], local[dst][cols[e]])                    */
  int64_t* nodes_len;               /* HOST [B*T] result, like samples_len                                          */
  /* ---- temporal filter (src/algo/neighbor_sampling.rs:36-77, src/python.rs:137-168); 0 = IdentityFilter ---- */
  int32_t filter_mode;              /* 0 none, 1 TEMPORAL_SAMPLE_STATIC, 2 _RELATIVE, 3 _DYNAMIC (reference mode + 1) */
  int32_t filter_forward;           /* FORWARD const generic (ignored for STATIC)                  */
  int64_t filter_window_lo;         /* inclusive window, RangeInclusive<i64>                       */
  int64_t filter_window_hi;
  const int64_t* const* timestamps; /* HOST [R] of DEVICE [nnz_r] i64 per CSC position             */
  const int64_t* const* inputs_state; /* HOST [T] of DEVICE [B, seeds_per_batch[t]] i64            */
  int64_t* const* states;           /* HOST [T] of DEVICE [B, samples_stride[t]] i64: per-sample
                                       filter state (scratch/output, caller allocated)             */
  /* ---- workspace ---- */
  void* workspace;                  /* DEVICE                                                      */
  size_t workspace_bytes;
  tchgeo_stream stream;
} tchgeo_sampling_args;

/* dst[i] = (int32) src[i]; TCHGEO_ERR_INDEX if a value is outside [0, 2^31).  Builds the optional
 * row_indices32 replica. */
TCHGEO_API tchgeo_status tchgeo_compress_indices(const int64_t* src /*DEVICE [n]*/, int64_t n,
                                                 int32_t* dst /*DEVICE [n]*/, int32_t* scratch /*DEVICE [1]*/,
                                                 tchgeo_stream stream);

/* Worst-case per-batch capacities (elements) for samples[t] and edges[r]:
 * the reference recurrence with every neighbourhood at full fanout. HOST outputs [T], [R]. */
TCHGEO_API tchgeo_status tchgeo_neighbor_sampling_capacity(const tchgeo_sampling_args* args, int64_t* samples_cap,
                                                int64_t* edges_cap);
TCHGEO_API size_t tchgeo_neighbor_sampling_workspace_bytes(const tchgeo_sampling_args* args);

/* replaces src/algo/neighbor_sampling.rs:233-356 (and :162-230 with T = R = 1).
 * If samples_len/edges_len/layer_offsets are all NULL the call is asynchronous: it returns after
 * enqueueing and device-side errors surface through tchgeo_neighbor_sampling_collect. */
TCHGEO_API tchgeo_status tchgeo_neighbor_sampling(const tchgeo_sampling_args* args);

/* Same as tchgeo_neighbor_sampling, but brackets every (hop, relation) kernel launch with CUDA events
 * on args->stream and returns the per-launch device durations (launch order = hop-major, relation
 * order within a hop; launches that cannot produce output are skipped).  Always synchronises.
 * launch_ms: HOST [launch_ms_cap]; num_launches: HOST [1]. Used by bench.py for the roofline figure. */
TCHGEO_API tchgeo_status tchgeo_neighbor_sampling_timed(const tchgeo_sampling_args* args, float* launch_ms,
                                                        int32_t launch_ms_cap, int32_t* num_launches);

/* Synchronise args->stream, fetch lengths / layer offsets / device error flag of the last
 * tchgeo_neighbor_sampling call that used args->workspace. */
TCHGEO_API tchgeo_status tchgeo_neighbor_sampling_collect(const tchgeo_sampling_args* args);

/* Flat-argument convenience wrapper for the homogeneous case (T = R = 1):
 * replaces src/algo/neighbor_sampling.rs:162-230 as called from src/python.rs:251.
 * layer_offsets: HOST [B*H*3]; out_lens: HOST [B*2] = (len(samples), len(edges)). */
TCHGEO_API tchgeo_status tchgeo_neighbor_sampling_homogenous(
    const int64_t* col_ptrs /*DEVICE [num_cols+1]*/, int64_t num_cols, const int64_t* row_indices /*DEVICE*/,
    const int64_t* inputs /*DEVICE [B,S]*/, int64_t num_batches, int64_t seeds_per_batch,
    const int64_t* num_neighbors /*HOST [H]*/, int32_t num_hops, int32_t sampler_kind,
    const double* weights /*DEVICE or NULL*/, uint64_t seed, uint32_t batch_base,
    int64_t* samples /*DEVICE [B,samples_stride]*/, int64_t samples_stride,
    int64_t* rows /*DEVICE [B,edges_stride]*/, int64_t* cols, int64_t* edge_index, int64_t edges_stride,
    int64_t* out_lens /*HOST or NULL*/, int64_t* layer_offsets /*HOST or NULL*/,
    void* workspace /*DEVICE*/, size_t workspace_bytes, tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* Sampling plan handle: everything a host would otherwise rebuild per call, kept behind the ABI.  A plan is a    */
/* deep copy of a tchgeo_sampling_args (so the caller's host arrays may go away), the launch plan derived from    */
/* it, pinned host memory for the length table and an event, so that a step is                                     */
/*     tchgeo_plan_enqueue(plan, seed, batch_base, stream)   H launches (+ the relabel stage), no host sync        */
/*     tchgeo_plan_collect(plan)                             waits for THAT enqueue only (an event, not the        */
/*                                                           stream), decodes lengths / layer offsets / errors    */
/* and two plans on two streams keep the device busy while the host reads the previous step's lengths.            */
/* The output buffers, the inputs buffers ([B, seeds_per_batch[t]], refilled by the caller before every enqueue,   */
/* e.g. with one H2D copy on the same stream) and the workspace named in `args` are BORROWED: caller-allocated     */
/* (tchgeo_neighbor_sampling_capacity / _workspace_bytes), they must outlive the plan.  args->samples_len /         */
/* edges_len / layer_offsets / nodes_len are ignored: results are read through tchgeo_plan_results.               */
/* replaces what src/python.rs:187-271 / :273-395 do around the algorithm call, for a host that samples repeatedly. */
/* -------------------------------------------------------------------------------------------- */
typedef struct tchgeo_plan tchgeo_plan_t;
TCHGEO_API tchgeo_status tchgeo_plan_create(const tchgeo_sampling_args* args, tchgeo_plan_t** out);
TCHGEO_API tchgeo_status tchgeo_plan_enqueue(tchgeo_plan_t* plan, uint64_t seed, uint32_t batch_base, tchgeo_stream stream);
/* Same, bracketing every hop launch (and the relabel stage as one interval) with CUDA events; synchronises and
 * collects.  launch_ms: HOST [cap]; the relabel interval, if any, comes last. */
TCHGEO_API tchgeo_status tchgeo_plan_enqueue_timed(tchgeo_plan_t* plan, uint64_t seed, uint32_t batch_base,
                                                   tchgeo_stream stream, float* launch_ms, int32_t launch_ms_cap,
                                                   int32_t* num_intervals);
TCHGEO_API tchgeo_status tchgeo_plan_collect(tchgeo_plan_t* plan);
/* HOST arrays owned by the plan, valid until the next collect: samples_len [B*T], edges_len [B*R],
 * layer_offsets [B*R*H*3], nodes_len [B*T] (NULL without relabel).  Any out pointer may be NULL. */
TCHGEO_API tchgeo_status tchgeo_plan_results(const tchgeo_plan_t* plan, const int64_t** samples_len,
                                             const int64_t** edges_len, const int64_t** layer_offsets,
                                             const int64_t** nodes_len);
/* Kernel launches one enqueue issues (hop kernels + fill + relabel kernels), for accounting. */
TCHGEO_API int32_t tchgeo_plan_num_launches(const tchgeo_plan_t* plan);
TCHGEO_API void tchgeo_plan_destroy(tchgeo_plan_t* plan);

/* -------------------------------------------------------------------------------------------- */
/* Range-partitioned CSC (graph split over ranks by column range; BASELINE config 5): owner side. */
/* Answers n requests for columns [col_begin, col_begin + ncols_local): request i = node id        */
/* req_ids[i], req_meta[i] = (batch << 32) | position of that node in its batch's samples vector.   */
/* out_ids / out_ptrs: [n, fanout], slot s of request i = sampled neighbour id and GLOBAL csc        */
/* position (edge_base + local position), -1 for unused slots.  Draws use the counters of the       */
/* replicated path, so a requester that lays the answers out in frontier order reproduces           */
/* tchgeo_neighbor_sampling bit for bit.  Synchronises the stream (error flag read-back).           */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API tchgeo_status tchgeo_serve_requests(
    const int64_t* ptrs_local /*DEVICE [ncols_local+1], rebased to indices_local*/,
    const int64_t* indices_local /*DEVICE*/, const double* weights_local /*DEVICE or NULL*/, int64_t col_begin,
    int64_t ncols_local, int64_t edge_base, const int64_t* req_ids /*DEVICE [n]*/, const int64_t* req_meta /*DEVICE [n]*/,
    int64_t n, int64_t fanout, int32_t sampler_kind, uint64_t seed, uint32_t rel, int64_t* out_ids /*DEVICE [n*fanout]*/,
    int64_t* out_ptrs /*DEVICE [n*fanout]*/, int32_t* err_scratch /*DEVICE [1]*/, tchgeo_stream stream);

/* Requester side of the partitioned path as a device pipeline.  A hop on every rank is
 *   tchgeo_part_begin_hop   count the frontier per owning rank (owner(w) = min(w / cols_per_rank, world-1)) into
 *                           counts[world] and scatter request rows (id, (batch_base+b) << 32 | pos) into `req`
 *                           grouped by owner (order inside a group is arbitrary); asynchronous
 *   -- exchange counts, then the request rows, with an all-to-all (NCCL; the host reads the counts once) --
 *   tchgeo_serve_requests_rows  owner side on interleaved rows: req [n,2] -> compact int32 ans [n, 2*fanout] = fanout
 *                           neighbour ids then fanout LOCAL CSC positions per row, -1 padded (half the exchange
 *                           bytes; needs node ids and a rank's CSC share below 2^31); asynchronous
 *   -- all-to-all of the answer rows back --
 *   tchgeo_part_finish_hop  per-node answer counts, exclusive scan in frontier order, then the tree layout of
 *                           src/algo/neighbor_sampling.rs:210-218 appended to the caller's [B, stride] buffers at
 *                           node_len_in / edge_len_in; writes the new lengths; asynchronous
 * The frontier of batch b is samples[b, fr_begin[b] .. fr_end[b]) (fr_begin NULL = 0); frontier_cap bounds its
 * size.  Device-side errors are OR-ed into *err_word (DEVICE u32, caller zeroes it; decode with
 * tchgeo_status_from_error_word). */
TCHGEO_API tchgeo_status tchgeo_part_begin_hop(const int64_t* samples /*DEVICE [B, samples_stride]*/, int64_t samples_stride,
                                               const int64_t* fr_begin /*DEVICE [B] or NULL*/,
                                               const int64_t* fr_end /*DEVICE [B]*/, int64_t num_batches,
                                               int64_t frontier_cap, int64_t cols_per_rank, int32_t world,
                                               uint32_t batch_base, int64_t* counts /*DEVICE [world]*/,
                                               int64_t* cursor /*DEVICE [world] scratch*/,
                                               int64_t* req /*DEVICE [B*frontier_cap, 2]*/, int32_t* err_word,
                                               void* workspace /*DEVICE, tchgeo_part_hop_workspace_bytes; the same
                                               buffer must be passed to tchgeo_part_finish_hop of this hop*/,
                                               size_t workspace_bytes, tchgeo_stream stream);
/* tchgeo_part_begin_hop in two halves, for the request exchange fused into the scatter kernel: count_hop fills
 * counts[world]; the ranks all-gather the [world, world] count matrix; scatter_hop then writes every request row into
 * the local `req`, grouped by owner (the layout kernel reads it) and, when peer_req != NULL, a second kernel copies group
 * o, a contiguous run of rows, into the owning rank's request buffer peer_req[o] (HOST array of `world` DEVICE pointers
 * into NVLink peer memory) from row peer_row0[o] on -- the rows the request all-to-all would have delivered it to
 * (peer_row0[o] = number of requests lower-ranked requesters send to owner o); whole lines per warp.  The caller synchronises the ranks before the owners
 * serve.  Both asynchronous; same workspace as tchgeo_part_finish_hop. */
TCHGEO_API tchgeo_status tchgeo_part_count_hop(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                               const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                               int64_t cols_per_rank, int32_t world, int64_t* counts /*DEVICE [world]*/,
                                               int64_t* cursor /*DEVICE [world] scratch*/, int32_t* err_word,
                                               void* workspace, size_t workspace_bytes, tchgeo_stream stream);
TCHGEO_API tchgeo_status tchgeo_part_scatter_hop(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                                 const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                                 int64_t cols_per_rank, int32_t world, uint32_t batch_base,
                                                 const int64_t* counts /*DEVICE [world], from count_hop*/,
                                                 int64_t* cursor /*DEVICE [world], zeroed by count_hop*/,
                                                 int64_t* req /*DEVICE [B*frontier_cap, 2]*/,
                                                 void* const* peer_req /*HOST [world] of DEVICE buffers, or NULL*/,
                                                 const int64_t* peer_row0 /*HOST [world]*/, int32_t* err_word,
                                                 void* workspace, size_t workspace_bytes, tchgeo_stream stream);
TCHGEO_API tchgeo_status tchgeo_serve_requests_rows(const int64_t* ptrs_local, const int64_t* indices_local,
                                                    const double* weights_local, int64_t col_begin, int64_t ncols_local,
                                                    int64_t nnz_local, const int64_t* req /*DEVICE [n,2]*/, int64_t n,
                                                    int64_t fanout, int32_t sampler_kind, uint64_t seed, uint32_t rel,
                                                    int32_t* ans /*DEVICE [n, 2*fanout] int32*/, int32_t* err_word,
                                                    tchgeo_stream stream);
/* Same owner side with the answer exchange fused into the kernel: the requests received from rank q are rows
 * [sum(recv_counts[:q]), +recv_counts[q]) of `req`, and the answer row of the i-th of them is stored directly into
 * rank q's answer buffer peer_ans[q] (DEVICE pointers into NVLink peer memory, e.g. the buffer_ptrs of a torch
 * symmetric-memory allocation; HOST array of `world` pointers) at row peer_row0[q] + i -- where the answer all-to-all
 * would have delivered it (peer_row0[q] = number of requests rank q sent to lower-ranked owners).  The caller
 * synchronises the ranks (a symmetric-memory barrier) before tchgeo_part_finish_hop reads its own buffer.
 * recv_counts / peer_row0: HOST [world].  Asynchronous. */
TCHGEO_API tchgeo_status tchgeo_serve_requests_rows_peer(const int64_t* ptrs_local, const int64_t* indices_local,
                                                         const double* weights_local, int64_t col_begin,
                                                         int64_t ncols_local, int64_t nnz_local,
                                                         const int64_t* req /*DEVICE [n,2]*/, int64_t n, int64_t fanout,
                                                         int32_t sampler_kind, uint64_t seed, uint32_t rel, int32_t world,
                                                         void* const* peer_ans /*HOST [world] of DEVICE int32 buffers*/,
                                                         const int64_t* recv_counts /*HOST [world]*/,
                                                         const int64_t* peer_row0 /*HOST [world]*/, int32_t* err_word,
                                                         tchgeo_stream stream);
TCHGEO_API size_t tchgeo_part_hop_workspace_bytes(int64_t num_batches, int64_t frontier_cap);
TCHGEO_API tchgeo_status tchgeo_part_finish_hop(const int64_t* req /*DEVICE [F,2] as sent*/, const int32_t* ans /*DEVICE [F, 2*fanout]*/,
                                                int64_t num_requests, int64_t fanout,
                                                const int64_t* owner_edge_base /*DEVICE [world]*/, int64_t cols_per_rank,
                                                int32_t world,
                                                const int64_t* fr_begin, const int64_t* fr_end, int64_t num_batches,
                                                int64_t frontier_cap,
                                                const int64_t* node_len_in /*DEVICE [B]*/, const int64_t* edge_len_in,
                                                int64_t* node_len_out, int64_t* edge_len_out, int64_t* samples,
                                                int64_t samples_stride, int64_t* rows, int64_t* cols, int64_t* edge_index,
                                                int64_t edges_stride, int32_t* err_word, void* workspace,
                                                size_t workspace_bytes, tchgeo_stream stream);
/* ---- the same path with BOTH exchanges and every count kept on the device ("fixed segments", csrc/partitioned_fixed.cu):
 * every (requester q, owner o) pair owns `seg_rows` rows in o's request buffer req_in [world, seg_rows, 2] (+ one entry of
 * o's count table cnt_in [world]) and in q's answer buffer ans_in [world, seg_rows, 2*fanout] (int32).  The buffers live in
 * NVLink peer memory (e.g. torch symmetric memory); peer_* are HOST arrays of `world` DEVICE pointers to every rank's
 * buffer.  A hop is: tchgeo_partf_scatter (frontier -> rows grouped by owner -> whole-line stores into the owners'
 * request segments + counts), a barrier over the ranks, tchgeo_partf_serve (the owner samples every segment it holds,
 * counts read on the device, and stores the answer rows at the same row index of the requester's answer segment), a
 * barrier, tchgeo_partf_finish (one pass: slot map -> answer row -> scan with decoupled look-back -> tree layout of
 * src/algo/neighbor_sampling.rs:210-218 and the new lengths).  No host synchronisation inside a hop; a segment that
 * overflows raises TCHGEO_ERR_CAPACITY through *err_word (DEVICE u32, caller-zeroed).  All asynchronous.
 * `me` = this rank; slot_of: DEVICE u32 [B * frontier_cap] scratch shared by scatter and finish of one hop; send: DEVICE
 * [world, seg_rows, 2] scratch; cursor: DEVICE [world] scratch.  Results equal tchgeo_neighbor_sampling bit for bit. */
TCHGEO_API size_t tchgeo_partf_workspace_bytes(int64_t num_batches, int64_t frontier_cap);
TCHGEO_API tchgeo_status tchgeo_partf_scatter(const int64_t* samples /*DEVICE [B, samples_stride]*/, int64_t samples_stride,
                                              const int64_t* fr_begin /*DEVICE [B] or NULL*/, const int64_t* fr_end /*DEVICE [B]*/,
                                              int64_t num_batches, int64_t frontier_cap, int64_t cols_per_rank, int32_t world,
                                              int32_t me, uint32_t batch_base, int64_t seg_rows, int64_t* send, int64_t* cursor,
                                              uint32_t* slot_of, void* const* peer_req /*HOST [world]*/,
                                              void* const* peer_cnt /*HOST [world]*/, int32_t* err_word, tchgeo_stream stream);
TCHGEO_API tchgeo_status tchgeo_partf_serve(const int64_t* ptrs_local, const int64_t* indices_local,
                                            const int32_t* indices32_local /*DEVICE or NULL: int32 replica*/,
                                            const double* weights_local, int64_t col_begin, int64_t ncols_local,
                                            int64_t nnz_local, const int64_t* req_in /*DEVICE [world, seg_rows, 2]*/,
                                            const int64_t* cnt_in /*DEVICE [world]*/, int64_t seg_rows,
                                            int64_t max_requests /*bound of cnt_in entries (grid size); 0 = seg_rows*/,
                                            int64_t fanout, int32_t sampler_kind, uint64_t seed, uint32_t rel, int32_t world,
                                            int32_t me, void* const* peer_ans /*HOST [world]*/, int32_t* err_word,
                                            tchgeo_stream stream);
TCHGEO_API tchgeo_status tchgeo_partf_finish(const int32_t* ans_in /*DEVICE [world, seg_rows, 2*fanout]*/,
                                             const uint32_t* slot_of, int64_t seg_rows, int64_t fanout,
                                             const int64_t* owner_edge_base /*DEVICE [world]*/, int32_t world,
                                             const int64_t* fr_begin, const int64_t* fr_end, int64_t num_batches,
                                             int64_t frontier_cap, const int64_t* node_len_in /*DEVICE [B]*/,
                                             const int64_t* edge_len_in, int64_t* node_len_out, int64_t* edge_len_out,
                                             int64_t* samples, int64_t samples_stride, int64_t* rows, int64_t* cols,
                                             int64_t* edge_index, int64_t edges_stride, int32_t* err_word, void* workspace,
                                             size_t workspace_bytes, tchgeo_stream stream);
/* Map a device error word (as accumulated by the asynchronous entry points) to a status + last-error string. */
TCHGEO_API tchgeo_status tchgeo_status_from_error_word(uint32_t word);

/* -------------------------------------------------------------------------------------------- */
/* node2vec random walk over CSR.  walks: [num_walks, walk_length+1] row-major, -1 padded.        */
/* Walker i draws with walker index walker_base + i (so sharded launches reproduce one big one).   */
/* replaces src/algo/random_walk.rs:10-75                                                        */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API tchgeo_status tchgeo_random_walk(const int64_t* row_ptrs /*DEVICE [num_rows+1]*/, int64_t num_rows,
                                 const int64_t* col_indices /*DEVICE*/, const int64_t* start /*DEVICE [num_walks]*/,
                                 int64_t num_walks, int64_t walk_length, float p, float q, uint64_t seed,
                                 int64_t walker_base, int64_t* walks /*DEVICE*/,
                                 int64_t* stats /*DEVICE [2] scratch: rejection-loop attempts, error flags*/,
                                 int64_t* attempts_out /*HOST [1] or NULL*/, tchgeo_stream stream);

/* Same, with an optional int32 replica of col_indices (tchgeo_compress_indices): the neighbour gathers
 * and the has_edge binary searches then read it instead (half the DRAM lines per adjacency). */
TCHGEO_API tchgeo_status tchgeo_random_walk_ex(const int64_t* row_ptrs, int64_t num_rows, const int64_t* col_indices,
                                               const int32_t* col_indices32 /*DEVICE or NULL*/, const int64_t* start,
                                               int64_t num_walks, int64_t walk_length, float p, float q, uint64_t seed,
                                               int64_t walker_base, int64_t* walks, int64_t* stats,
                                               int64_t* attempts_out, tchgeo_stream stream);

/* Same over relation `rel` of a graph handle holding the CSR (row_ptrs / col_indices); uses the handle's int32 replica. */
TCHGEO_API tchgeo_status tchgeo_random_walk_graph(const tchgeo_graph_t* graph, int32_t rel, const int64_t* start,
                                                  int64_t num_walks, int64_t walk_length, float p, float q, uint64_t seed,
                                                  int64_t walker_base, int64_t* walks, int64_t* stats,
                                                  int64_t* attempts_out, tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* Negative neighbour sampling over CSR.  For every input v of every node type    */
/* and each of num_neg slots: pick one of the relations that start at v's type (uniformly, only     */
/* when there are several), draw up to try_count candidates w uniform in [0, node_count[rel]) and   */
/* keep the first with !has_edge(v, w) (inbound: !has_edge(w, v)) and v != w.  samples[t] = inputs[t]*/
/* ++ accepted candidates of dst type t at first appearance (insertion-order HashMap semantic, a    */
/* duplicated input maps to its last position); edges of relation r in generation order: rows = index*/
/* of the input, cols = local id of the candidate.  Relations are visited in array order and node   */
/* types in index order (the reference iterates HashMaps).                                          */
/* replaces src/algo/negative_sampling.rs:6-47 and :49-131, called from src/python.rs:689-783        */
/* -------------------------------------------------------------------------------------------- */
typedef struct tchgeo_negative_args {
  int32_t num_node_types;            /* T (1 for homogeneous)                                      */
  int32_t num_rels;                  /* R (1 for homogeneous)                                      */
  const int32_t* rel_src;            /* HOST [R]                                                   */
  const int32_t* rel_dst;            /* HOST [R]                                                   */
  const int64_t* const* row_ptrs;    /* HOST [R] of DEVICE [num_rows[r]+1] (CSR)                    */
  const int64_t* const* col_indices; /* HOST [R] of DEVICE [nnz_r], ascending inside a row          */
  const int64_t* num_rows;           /* HOST [R] rows of the CSR                                   */
  const int64_t* node_count;         /* HOST [R] size.1: candidates come from [0, node_count[r])    */
  const int64_t* const* inputs;      /* HOST [T] of DEVICE [num_inputs[t]]                          */
  const int64_t* num_inputs;         /* HOST [T] (0 = type absent from inputs)                      */
  int64_t num_neg;
  int64_t try_count;
  int32_t inbound;                   /* heterogenous only: test has_edge(w, v) instead of (v, w)    */
  int32_t reserved0;
  uint64_t seed;
  int64_t* const* samples;           /* HOST [T] of DEVICE [samples_cap[t]]                         */
  int64_t* const* rows;              /* HOST [R] of DEVICE [edges_cap[r]]                           */
  int64_t* const* cols;              /* HOST [R]                                                   */
  int64_t* samples_len;              /* HOST [T] out                                               */
  int64_t* edges_len;                /* HOST [R] out                                               */
  void* workspace;                   /* DEVICE                                                     */
  size_t workspace_bytes;
  tchgeo_stream stream;
} tchgeo_negative_args;

TCHGEO_API tchgeo_status tchgeo_negative_sampling_capacity(const tchgeo_negative_args* args, int64_t* samples_cap /*HOST [T]*/,
                                                           int64_t* edges_cap /*HOST [R]*/);
TCHGEO_API size_t tchgeo_negative_sampling_workspace_bytes(const tchgeo_negative_args* args);
/* Synchronises args->stream (lengths are returned to the host). */
TCHGEO_API tchgeo_status tchgeo_negative_sampling(const tchgeo_negative_args* args);

/* Temporal random walk.  walks / walks_timestamps: [num_walks, walk_length] (NOT +1), column 0 =
 * start / start timestamp.  Per step the next node is drawn with the reference's k = 1 reservoir among the
 * neighbours whose timestamp (edge timestamp, or the neighbour's node timestamp when that is -1) is -1 or lies in
 * [start_ts + window_lo, start_ts + window_hi) (everything passes when start_ts is -1); a step without such a
 * neighbour restarts from a uniformly drawn earlier position of the same walk.
 * replaces src/algo/random_walk.rs:80-158, called from src/python.rs:610-643 */
TCHGEO_API tchgeo_status tchgeo_tempo_random_walk(const int64_t* row_ptrs /*DEVICE [num_rows+1]*/, int64_t num_rows,
                                                  const int64_t* col_indices /*DEVICE [nnz]*/,
                                                  const int64_t* node_timestamps /*DEVICE [num_node_timestamps]*/,
                                                  int64_t num_node_timestamps,
                                                  const int64_t* edge_timestamps /*DEVICE [nnz]*/,
                                                  const int64_t* start /*DEVICE [num_walks]*/,
                                                  const int64_t* start_timestamps /*DEVICE [num_walks]*/, int64_t num_walks,
                                                  int64_t walk_length, int64_t window_lo, int64_t window_hi, uint64_t seed,
                                                  int64_t walker_base, int64_t* walks, int64_t* walks_timestamps,
                                                  int32_t* scratch /*DEVICE [1]*/, tchgeo_stream stream);

/* -------------------------------------------------------------------------------------------- */
/* Row gather: dst[i, :] = src[index[i], :] for rows of row_bytes bytes of any     */
/* dtype.  The step after the sampler in every loader (x[samples], edge_attr[perm[edge_index]];      */
/* examples/neighbor_sampling.py:21-24).  TCHGEO_ERR_INDEX for an index outside [0, num_rows); with */
/* scratch == NULL the call is asynchronous and does not validate (bad rows are left unwritten).     */
/* -------------------------------------------------------------------------------------------- */
TCHGEO_API tchgeo_status tchgeo_gather_rows(const void* src /*DEVICE [num_rows, row_bytes]*/, int64_t num_rows,
                                            int64_t row_bytes, const int64_t* index /*DEVICE [n]*/, int64_t n,
                                            void* dst /*DEVICE [n, row_bytes]*/, int32_t* scratch /*DEVICE [1] or NULL*/,
                                            tchgeo_stream stream);

/* Ragged pack: dst[off[b] + i] = src[b*stride + i] for i < min(lens[b*lens_stride], max_len), off = exclusive prefix
 * sum of the clamped lengths, written to offsets[0..B] (offsets[B] = total).  The sampler's outputs live in padded
 * [B, capacity] buffers; a host consumer packs the used prefixes and copies them back with ONE transfer per tensor
 * (what src/python.rs:259-262 does per call with Vec -> Tensor copies).  lens / offsets: DEVICE.  Asynchronous. */
TCHGEO_API tchgeo_status tchgeo_pack_ragged(const int64_t* src /*DEVICE [B, stride]*/, int64_t stride,
                                            const int64_t* lens /*DEVICE*/, int64_t lens_stride, int64_t num_batches,
                                            int64_t max_len, int64_t* dst /*DEVICE [sum]*/, int64_t* offsets /*DEVICE [B+1]*/,
                                            tchgeo_stream stream);

/* Compact host transport of a group of homogeneous batches (extension; the reference builds its output Vecs in host
 * memory to begin with, src/python.rs:259-262, so it has no counterpart to cite).  The reference layout is 24 B per
 * sampled edge on the bus (samples, cols, edge_index as i64; rows is an arange, src/algo/neighbor_sampling.rs:210-218).
 * tchgeo_pack_transport (DEVICE, asynchronous) writes the used prefixes of `count` padded batches back to back as
 *   samples32 [sum n_b] i32, eidx32 [sum e_b] i32, counts [sum n_b] u8 = edges drawn for node j of the batch, i.e. the
 *   run length of j in the batch's (non-decreasing) cols vector,
 * and the packed offsets n_off / e_off (DEVICE [count + 1]); counts_bytes = size of `counts` (zeroed here).  An id that
 * does not fit i32 raises TCHGEO_ERR_INDEX in *err_word, a run longer than 255 or a cols vector that is not
 * non-decreasing TCHGEO_ERR_CAPACITY: the caller then uses tchgeo_pack_ragged.
 * tchgeo_host_unpack_transport (HOST memory only, synchronous, num_threads worker threads, non-temporal stores) rebuilds
 * samples / cols / edge_index [same packed offsets] as i64 -- byte for byte what tchgeo_pack_ragged + a D2H copy land.
 * eidx32 (and, on the host side, edge_index) may be NULL in both calls: edge_index then travels as i64 through
 * tchgeo_pack_ragged -- 13 B per edge on the bus and a third less for the host threads to write, the better balance
 * on a host with few cores. */
TCHGEO_API tchgeo_status tchgeo_pack_transport(const int64_t* samples /*DEVICE [B, samples_stride]*/, int64_t samples_stride,
                                               const int64_t* cols /*DEVICE [B, edges_stride]*/,
                                               const int64_t* edge_index /*DEVICE [B, edges_stride]*/, int64_t edges_stride,
                                               const int64_t* n_lens /*DEVICE [count]*/, const int64_t* e_lens /*DEVICE [count]*/,
                                               int64_t count, int64_t max_n, int64_t max_e, int32_t* samples32 /*DEVICE*/,
                                               int32_t* eidx32 /*DEVICE*/, uint8_t* counts /*DEVICE*/, int64_t counts_bytes,
                                               int64_t* n_off /*DEVICE [count+1]*/, int64_t* e_off /*DEVICE [count+1]*/,
                                               int32_t* err_word /*DEVICE*/, tchgeo_stream stream);
TCHGEO_API tchgeo_status tchgeo_host_unpack_transport(const int32_t* samples32 /*HOST*/, const int32_t* eidx32 /*HOST*/,
                                                      const uint8_t* counts /*HOST*/, const int64_t* n_off /*HOST [count+1]*/,
                                                      const int64_t* e_off /*HOST [count+1]*/, int64_t count,
                                                      int64_t* samples /*HOST*/, int64_t* cols /*HOST*/,
                                                      int64_t* edge_index /*HOST*/, int32_t num_threads);

/* -------------------------------------------------------------------------------------------- */
/* Dedup + insertion-order relabel of one sampled tree (additive stage).                          */
/*   nodes      = seeds (duplicates kept) ++ every other id at first appearance                    */
/*   local[i]   = index into `nodes` of samples[i] (a duplicated seed maps to its LAST seed slot)  */
/* semantic of src/algo/negative_sampling.rs:20-47                                               */
/* -------------------------------------------------------------------------------------------- */
/* Batched form: B trees in padded [B, stride] rows, tree b holding lens[b] <= n_max ids (DEVICE lengths: no host
 * round trip), the first num_seeds of them seeds.  nodes / local: [B, stride]; nodes_len: DEVICE [B].
 * id_bound states what the caller knows about the ids:
 *   0               any non-negative i64 (global hash tables with 16-byte slots, processed in L2-sized waves);
 *   1 .. 2^32-1     every id is in [0, id_bound); an id outside raises TCHGEO_ERR_INDEX.  Pass the node count when it is
 *                   known (the sampling plan does, from the graph handle): bounds up to 2^25 select direct-address
 *                   shared-memory tables, larger ones (2^32-1 = "fits 32 bits, nothing else known") hashed ones.
 * Asynchronous; device-side errors are OR-ed into *err_word (DEVICE u32, caller-zeroed; tchgeo_status_from_error_word).
 * The workspace size depends on id_bound: pass the same value to the query and to the call (see csrc/relabel.cu). */
TCHGEO_API size_t tchgeo_unique_relabel_batched_workspace_bytes(int64_t num_batches, int64_t n_max, int64_t id_bound);
TCHGEO_API tchgeo_status tchgeo_unique_relabel_batched(const int64_t* samples /*DEVICE [B, stride]*/, int64_t stride,
                                                       const int64_t* lens /*DEVICE [B]*/, int64_t num_batches,
                                                       int64_t num_seeds, int64_t n_max, int64_t id_bound,
                                                       int64_t* nodes /*DEVICE [B, stride]*/, int64_t* local /*DEVICE [B, stride]*/,
                                                       int64_t* nodes_len /*DEVICE [B]*/, void* workspace /*DEVICE*/,
                                                       size_t workspace_bytes, int32_t* err_word /*DEVICE*/,
                                                       tchgeo_stream stream);
TCHGEO_API size_t tchgeo_unique_relabel_workspace_bytes(int64_t n);
TCHGEO_API tchgeo_status tchgeo_unique_relabel(const int64_t* samples /*DEVICE [n]*/, int64_t n, int64_t num_seeds,
                                    int64_t* nodes /*DEVICE [n]*/, int64_t* local /*DEVICE [n]*/,
                                    int64_t* num_nodes /*HOST [1]*/, void* workspace /*DEVICE*/,
                                    size_t workspace_bytes, tchgeo_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* TCHGEO_CUDA_H_ */
