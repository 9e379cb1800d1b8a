// Negative neighbour sampling on sm_100a.
//
// Replaces src/algo/negative_sampling.rs:6-47 (homogenous) and :49-131 (heterogenous) of the reference, called
// from src/python.rs:689-783.  For every input node v and each of `num_neg` slots the reference draws up to
// `try_count` candidates w uniform in [0, size.1) and keeps the first with !has_edge(v, w) && v != w; accepted
// candidates are appended to the dst type's sample list through a HashMap (inputs first, a duplicated input
// maps to its LAST position, new ids at first appearance) and the edge (i, local id of w) is recorded.
//
// Parallel formulation
//   1. neg_draw_kernel: one thread per (input i, slot): relation choice (heterogeneous only) and the try loop,
//      every draw from Philox(seed; i, slot, attempt / 4) so results do not depend on launch geometry;
//      has_edge is the binary search of graph.rs:80-83 over the row's sorted col_indices.
//   2. per dst type: stable compaction (CUB exclusive scan) of the accepted candidates in the reference's
//      generation order (node type, input, slot), then the insertion-order relabel of relabel.cu
//      (tchgeo_unique_relabel reproduces the HashMap exactly).
//   3. per relation: stable compaction of its accepted slots -> rows = i, cols = local id.
// HBM-bound integer work (random row_ptrs / col_indices probes); no tensor cores.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <vector>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace tchgeo {
// csrc/relabel.cu: the batched dedup + relabel stage (device-side lengths, asynchronous)
size_t relabel_workspace_bytes(int64_t num_trees, int64_t n_max, int64_t id_bound, bool prefer_waves);
bool relabel_is_bucketed(int64_t num_trees, int64_t n_max, int64_t id_bound);
tchgeo_status relabel_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees,
                              int64_t num_seeds, int64_t n_max, int64_t id_bound, int64_t* nodes, int64_t* local,
                              int64_t* nodes_len, void* workspace, size_t workspace_bytes, uint32_t* err,
                              cudaStream_t stream, bool prefer_waves);
namespace {

constexpr int NEG_THREADS = 256;
constexpr uint32_t TAG_NEGATIVE = 5u;
constexpr uint32_t NEG_BLOCK_REL = 0xFFFFFFFFu;  // counter block of the relation-choice draw

struct NegRel {  // one candidate relation of the current source node type
  const int64_t* row_ptrs;
  const int64_t* col_indices;
  int64_t num_rows;   // rows of the CSR (row_ptrs has num_rows + 1 entries)
  int64_t node_count; // size.1: candidates are drawn from [0, node_count)
  int32_t rel;        // global relation index
  int32_t pad;
};

struct NegDrawParams {
  const int64_t* inputs;
  int64_t num_inputs, num_neg, try_count;
  const NegRel* rels;  // DEVICE [n_choices]
  int32_t n_choices;
  int32_t choose;      // 1: draw the relation among n_choices > 1, 0: rels[0]
  int32_t inbound;
  uint32_t key0, key1, tag;
  int64_t* cand;       // [num_inputs * num_neg] accepted candidate or -1
  int32_t* crel;       // [num_inputs * num_neg] global relation index of the slot
  uint32_t* err;
};

__device__ __forceinline__ bool neg_has_edge(const NegRel& g, int64_t x, int64_t y, uint32_t* err) {
  if (x < 0 || x >= g.num_rows) {  // ptrs[x + 1] out of bounds: the reference panics
    atomicOr(err, DEV_ERR_INDEX);
    return true;
  }
  int64_t lo = __ldg(g.row_ptrs + x), hi = __ldg(g.row_ptrs + x + 1);
  while (lo < hi) {  // graph.rs:80-83
    const int64_t mid = lo + ((hi - lo) >> 1);
    const int64_t v = __ldg(g.col_indices + mid);
    if (v == y) return true;
    if (v < y) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(NEG_THREADS) neg_draw_kernel(const NegDrawParams p) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.num_inputs * p.num_neg) return;
  const int64_t i = g / p.num_neg;
  const uint32_t slot = (uint32_t)(g - i * p.num_neg);
  const int64_t v = p.inputs[i];
  uint32_t c = 0;
  if (p.choose) {  // negative_sampling.rs:104: rng.gen_range(0..node_rels.len())
    const Philox4 r = philox4x32_10((uint32_t)i, slot, NEG_BLOCK_REL, p.tag, p.key0, p.key1);
    c = __umulhi(r.x, (uint32_t)p.n_choices);
  }
  const NegRel rel = p.rels[c];
  int64_t found = -1;
  Philox4 r4 = {0, 0, 0, 0};
  for (int64_t t = 0; t < p.try_count; ++t) {  // :33-43 / :110-126
    if ((t & 3) == 0) r4 = philox4x32_10((uint32_t)i, slot, (uint32_t)(t >> 2), p.tag, p.key0, p.key1);
    const int64_t w = (int64_t)__umulhi(pick4(r4, (uint32_t)(t & 3)), (uint32_t)rel.node_count);
    const bool he = p.inbound ? neg_has_edge(rel, w, v, p.err) : neg_has_edge(rel, v, w, p.err);
    if (!he && v != w) {
      found = w;
      break;
    }
  }
  p.cand[g] = found;
  p.crel[g] = rel.rel;
}

// flag(g) = slot g was accepted and belongs to dst type `t` (by_type) / relation `t` (!by_type).  The scan reads it through
// a transform iterator and the scatter kernels recompute it: no flag array, no flag kernel.
struct NegFlag {
  const int64_t* cand;
  const int32_t* crel;
  const int32_t* rel_dst;
  int t, by_type;
  __host__ __device__ __forceinline__ int operator()(int64_t g) const {
    const int r = crel[g];
    return (cand[g] >= 0 && (by_type ? rel_dst[r] : r) == t) ? 1 : 0;
  }
};

__global__ void __launch_bounds__(NEG_THREADS) neg_scatter_type_kernel(const int64_t* __restrict__ cand, NegFlag flag,
                                                                      const int* __restrict__ ranks, int64_t G,
                                                                      int64_t* __restrict__ seq_tail,
                                                                      int* __restrict__ tpos, int64_t* total) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int f = flag(g);
  if (f) {
    seq_tail[ranks[g]] = cand[g];
    tpos[g] = ranks[g];
  }
  if (g == G - 1) *total = (int64_t)ranks[g] + f;
}

// rows[e] = index of the input inside its type's inputs, cols[e] = local id of the accepted candidate
__global__ void __launch_bounds__(NEG_THREADS) neg_scatter_rel_kernel(NegFlag flag, const int* __restrict__ ranks,
                                                                     const int* __restrict__ tpos, int64_t g0,
                                                                     int64_t g1, int64_t G, int64_t num_neg,
                                                                     const int64_t* __restrict__ local_tail,
                                                                     int64_t* __restrict__ rows,
                                                                     int64_t* __restrict__ cols, int64_t* total) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int f = flag(g);
  if (f && g >= g0 && g < g1) {
    const int e = ranks[g];
    rows[e] = (g - g0) / num_neg;
    cols[e] = local_tail[tpos[g]];
  }
  if (g == G - 1) *total = (int64_t)ranks[g] + f;
}

using FlagIterator = cub::TransformInputIterator<int, NegFlag, cub::CountingInputIterator<int64_t>>;
inline FlagIterator flag_iterator(const NegFlag& f) { return FlagIterator(cub::CountingInputIterator<int64_t>(0), f); }

// len_out = num_seeds + accepted (accepted == NULL: nothing can be appended)
__global__ void neg_len_kernel(const int64_t* accepted, int64_t num_seeds, int64_t* len_out) {
  *len_out = num_seeds + (accepted ? *accepted : 0);
}

inline size_t up(size_t x) { return (x + 255) / 256 * 256; }

struct NegPlan {
  int T = 0, R = 0;
  int64_t G = 0;                    // all slots, input types concatenated in node-type order
  std::vector<int64_t> S;           // inputs per type (0 when absent)
  std::vector<int64_t> slot_base;   // first slot of each input type
  std::vector<int64_t> seq_cap;     // S[t] + number of slots that can target type t
  std::vector<int64_t> seq_off;     // offset of type t's segment in seq / local
  int64_t seq_total = 0, max_seq = 0;
  size_t cub_bytes = 0, rl_bytes = 0;
  std::vector<int64_t> id_bound;    // [T] what the relabel stage may assume about the type's ids (0: any i64)
  bool prefer_waves = true;
  size_t off_cnts = 0;              // device counters: accepted[T] | seq_len[T] | nodes_len[T] | edges_len[R]
  size_t off_hdr, off_rels, off_rel_dst, off_cand, off_crel, off_tpos, off_ranks, off_seq, off_local, off_cub,
      off_rl, total;
};

tchgeo_status neg_plan(const tchgeo_negative_args* a, NegPlan& P, bool layout = true) {
  TCHGEO_REQUIRE(a, "NULL args");
  const int T = a->num_node_types, R = a->num_rels;
  TCHGEO_REQUIRE(T >= 1 && R >= 1, "need at least one node type and one relation");
  TCHGEO_REQUIRE(a->rel_src && a->rel_dst && a->row_ptrs && a->col_indices && a->num_rows && a->node_count &&
                     a->inputs && a->num_inputs,
                 "NULL table");
  TCHGEO_REQUIRE(a->num_neg >= 0 && a->try_count >= 0, "negative num_neg / try_count");
  P.T = T; P.R = R;
  P.S.assign(T, 0); P.slot_base.assign(T, 0); P.seq_cap.assign(T, 0); P.seq_off.assign(T, 0);
  for (int r = 0; r < R; ++r) {
    TCHGEO_REQUIRE(a->rel_src[r] >= 0 && a->rel_src[r] < T && a->rel_dst[r] >= 0 && a->rel_dst[r] < T,
                   "relation %d: node type out of range", r);
    TCHGEO_REQUIRE(a->node_count[r] >= 0 && a->node_count[r] < ((int64_t)1 << 32), "relation %d: size out of range", r);
  }
  // ids of type t: its inputs (checked against num_rows by the draw kernel when the type is the source of a relation)
  // and the candidates drawn from [0, node_count) of the relations it is the destination of.  With such a bound the
  // relabel stage can use 32-bit keys; a type whose inputs nothing checks keeps 64-bit keys.  The call relabels ONE tree
  // of millions of ids per node type, for which the wave form is the fast one (1 Mi inputs x 5 negatives: waves with
  // 64-bit keys 1.12 ms per call, with 32-bit keys 0.96 ms, direct-address buckets -- built for hundreds of trees --
  // 1.3 x the wave time).  TCHGEO_NEG_RELABEL=direct | waves64 select the other two.
  P.id_bound.assign(T, 0);
  {
    const char* f = getenv("TCHGEO_NEG_RELABEL");
    P.prefer_waves = !(f && strcmp(f, "direct") == 0);
    const bool waves64 = f && strcmp(f, "waves64") == 0;
    for (int t = 0; t < T && !waves64; ++t) {
      int64_t bound = 0;
      bool is_src = false;
      for (int r = 0; r < R; ++r) {
        if (a->rel_src[r] == t) { is_src = true; bound = std::max(bound, a->num_rows[r]); }
        if (a->rel_dst[r] == t) bound = std::max(bound, a->node_count[r]);
      }
      const bool has_inputs = a->num_inputs[t] > 0;
      if (bound > 0 && bound < 0xFFFFFFFFll && (is_src || !has_inputs)) P.id_bound[t] = bound;
    }
  }
  int64_t G = 0;
  for (int t = 0; t < T; ++t) {
    const int64_t s = a->num_inputs[t] > 0 ? a->num_inputs[t] : 0;
    TCHGEO_REQUIRE(s < ((int64_t)1 << 32), "too many inputs");
    TCHGEO_REQUIRE(s == 0 || a->inputs[t], "inputs[%d] is NULL", t);
    P.S[t] = s;
    P.slot_base[t] = G;
    G += s * a->num_neg;
  }
  TCHGEO_REQUIRE(G < ((int64_t)1 << 30), "too many (input, slot) pairs for one call");
  P.G = G;
  for (int t = 0; t < T; ++t) {
    int64_t cap = P.S[t];
    for (int s = 0; s < T; ++s) {
      bool feeds = false;
      for (int r = 0; r < R; ++r) feeds |= (a->rel_src[r] == s && a->rel_dst[r] == t);
      if (feeds) cap += P.S[s] * a->num_neg;
    }
    P.seq_cap[t] = cap;
    P.seq_off[t] = P.seq_total;
    P.seq_total += cap;
    if (cap > P.max_seq) P.max_seq = cap;
  }
  TCHGEO_REQUIRE(P.max_seq < ((int64_t)1 << 30), "sample list too long for one call");
  if (!layout) return TCHGEO_OK;  // capacities only: pure host arithmetic (the CUB size queries need a device)
  size_t cub = 0;
  TCHGEO_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, cub, flag_iterator(NegFlag{nullptr, nullptr, nullptr, 0, 0}),
                                                  (int*)nullptr, (int64_t)(G > 0 ? G : 1)));
  P.cub_bytes = cub;
  P.rl_bytes = 0;
  for (int t = 0; t < T; ++t) {
    if (!P.prefer_waves && !relabel_is_bucketed(1, std::max<int64_t>(P.seq_cap[t], 1), P.id_bound[t])) P.id_bound[t] = 0;
    const size_t need = relabel_workspace_bytes(1, std::max<int64_t>(P.seq_cap[t], 1), P.id_bound[t], P.prefer_waves);
    TCHGEO_REQUIRE(need != 0, "relabel workspace query failed");
    P.rl_bytes = std::max(P.rl_bytes, need);
  }
  const size_t g = (size_t)(G > 0 ? G : 1), sq = (size_t)(P.seq_total > 0 ? P.seq_total : 1);
  size_t o = 0;
  P.off_hdr = o; o += 256;  // [8] err (u32)
  P.off_cnts = o; o += up((size_t)(3 * T + R) * 8);
  P.off_rels = o; o += up((size_t)R * sizeof(NegRel));
  P.off_rel_dst = o; o += up((size_t)R * 4);
  P.off_cand = o; o += up(g * 8);
  P.off_crel = o; o += up(g * 4);
  P.off_tpos = o; o += up(g * 4);
  P.off_ranks = o; o += up(g * 4);
  P.off_seq = o; o += up(sq * 8);
  P.off_local = o; o += up(sq * 8);
  P.off_cub = o; o += up(cub);
  P.off_rl = o; o += up(P.rl_bytes);
  P.total = o + 256;
  return TCHGEO_OK;
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" size_t tchgeo_negative_sampling_workspace_bytes(const tchgeo_negative_args* args) {
  NegPlan P;
  if (neg_plan(args, P) != TCHGEO_OK) return 0;
  return P.total;
}

extern "C" tchgeo_status tchgeo_negative_sampling_capacity(const tchgeo_negative_args* args, int64_t* samples_cap,
                                                           int64_t* edges_cap) {
  NegPlan P;
  const tchgeo_status st = neg_plan(args, P, false);
  if (st != TCHGEO_OK) return st;
  TCHGEO_REQUIRE(samples_cap && edges_cap, "NULL output");
  for (int t = 0; t < P.T; ++t) samples_cap[t] = P.seq_cap[t];
  for (int r = 0; r < P.R; ++r) edges_cap[r] = P.S[args->rel_src[r]] * args->num_neg;
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_negative_sampling(const tchgeo_negative_args* a) {
  NegPlan P;
  tchgeo_status st = neg_plan(a, P);
  if (st != TCHGEO_OK) return st;
  TCHGEO_REQUIRE(a->samples && a->rows && a->cols && a->samples_len && a->edges_len, "NULL output table");
  TCHGEO_REQUIRE(a->workspace && a->workspace_bytes >= P.total, "workspace too small: need %zu bytes", P.total);
  const int T = P.T, R = P.R;
  cudaStream_t stream = (cudaStream_t)a->stream;
  char* ws = (char*)a->workspace;
  uint32_t* d_err = (uint32_t*)(ws + P.off_hdr + 8);
  NegRel* d_rels = (NegRel*)(ws + P.off_rels);
  int32_t* d_rel_dst = (int32_t*)(ws + P.off_rel_dst);
  int64_t* cand = (int64_t*)(ws + P.off_cand);
  int32_t* crel = (int32_t*)(ws + P.off_crel);
  int* tpos = (int*)(ws + P.off_tpos);
  int* ranks = (int*)(ws + P.off_ranks);
  int64_t* seq = (int64_t*)(ws + P.off_seq);
  int64_t* local = (int64_t*)(ws + P.off_local);
  for (int t = 0; t < T; ++t) a->samples_len[t] = 0;
  for (int r = 0; r < R; ++r) a->edges_len[r] = 0;

  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + P.off_hdr, 0, 256, stream));
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(d_rel_dst, a->rel_dst, (size_t)R * 4, cudaMemcpyHostToDevice, stream));
  // relation tables grouped by source type, in relation order (node_rels, negative_sampling.rs:66-73)
  std::vector<NegRel> h_rels;
  std::vector<int> first(T, 0), count(T, 0);
  for (int s = 0; s < T; ++s) {
    first[s] = (int)h_rels.size();
    for (int r = 0; r < R; ++r) {
      if (a->rel_src[r] != s) continue;
      NegRel nr;
      nr.row_ptrs = a->row_ptrs[r]; nr.col_indices = a->col_indices[r];
      nr.num_rows = a->num_rows[r]; nr.node_count = a->node_count[r];
      nr.rel = r; nr.pad = 0;
      h_rels.push_back(nr);
      ++count[s];
    }
  }
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(d_rels, h_rels.data(), h_rels.size() * sizeof(NegRel), cudaMemcpyHostToDevice, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));  // h_rels / a->rel_dst are pageable host memory

  // ---- 1. draws ---------------------------------------------------------------------------------
  for (int s = 0; s < T; ++s) {
    const int64_t slots = P.S[s] * a->num_neg;
    if (slots == 0) continue;
    if (count[s] == 0) {  // &node_rels[node_type] on a missing key (:98) panics
      set_last_error("node type %d has inputs but no relation starts from it (the reference panics)", s);
      return TCHGEO_ERR_REFERENCE_PANIC;
    }
    for (int c = 0; c < count[s]; ++c) {
      const NegRel& nr = h_rels[first[s] + c];
      TCHGEO_REQUIRE(nr.row_ptrs && (nr.col_indices || nr.num_rows == 0), "relation %d: NULL graph", nr.rel);
      if (nr.node_count == 0 && a->try_count > 0) {  // gen_range(0..0) panics
        set_last_error("relation %d has no destination nodes to draw from (the reference panics)", nr.rel);
        return TCHGEO_ERR_REFERENCE_PANIC;
      }
    }
    NegDrawParams dp;
    dp.inputs = a->inputs[s]; dp.num_inputs = P.S[s]; dp.num_neg = a->num_neg; dp.try_count = a->try_count;
    dp.rels = d_rels + first[s]; dp.n_choices = count[s]; dp.choose = count[s] > 1 ? 1 : 0;  // with one candidate relation the draw always yields it
    dp.inbound = a->inbound ? 1 : 0;
    dp.key0 = (uint32_t)a->seed; dp.key1 = (uint32_t)(a->seed >> 32);
    dp.tag = TAG_NEGATIVE | ((uint32_t)s << 8);
    dp.cand = cand + P.slot_base[s]; dp.crel = crel + P.slot_base[s]; dp.err = d_err;
    neg_draw_kernel<<<(unsigned)((slots + NEG_THREADS - 1) / NEG_THREADS), NEG_THREADS, 0, stream>>>(dp);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }

  const int64_t G = P.G;
  const unsigned ggrid = (unsigned)((G + NEG_THREADS - 1) / NEG_THREADS);
  // every count stays on the device until the single read-back at the end of the call (round 1 synchronised once per
  // node type, once more inside the relabel stage and once per relation)
  int64_t* d_acc = (int64_t*)(ws + P.off_cnts);   // accepted candidates per dst type
  int64_t* d_len = d_acc + T;                     // inputs + accepted: the length the relabel stage reads
  int64_t* d_nodes = d_len + T;                   // distinct nodes per type
  int64_t* d_elen = d_nodes + T;                  // edges per relation
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(d_acc, 0, (size_t)(3 * T + R) * 8, stream));
  // ---- 2. per dst type: sequence = inputs ++ accepted candidates, relabel -----------------------------
  for (int t = 0; t < T; ++t) {
    int64_t* seq_t = seq + P.seq_off[t];
    int64_t* local_t = local + P.seq_off[t];
    if (P.S[t] > 0)
      TCHGEO_CUDA_CHECK(cudaMemcpyAsync(seq_t, a->inputs[t], (size_t)P.S[t] * 8, cudaMemcpyDeviceToDevice, stream));
    const bool grows = G > 0 && P.seq_cap[t] > P.S[t];
    if (grows) {
      const NegFlag flag{cand, crel, d_rel_dst, t, 1};
      size_t cub = P.cub_bytes;
      TCHGEO_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws + P.off_cub, cub, flag_iterator(flag), ranks, G, stream));
      neg_scatter_type_kernel<<<ggrid, NEG_THREADS, 0, stream>>>(cand, flag, ranks, G, seq_t + P.S[t], tpos, d_acc + t);
      TCHGEO_CUDA_CHECK(cudaGetLastError());
    }
    neg_len_kernel<<<1, 1, 0, stream>>>(grows ? d_acc + t : nullptr, P.S[t], d_len + t);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    if (P.seq_cap[t] == 0) continue;
    TCHGEO_REQUIRE(a->samples[t], "samples[%d] is NULL", t);
    st = relabel_enqueue(seq_t, P.seq_cap[t], d_len + t, 1, P.S[t], P.seq_cap[t], P.id_bound[t], a->samples[t], local_t, d_nodes + t,
                         ws + P.off_rl, P.rl_bytes, d_err, stream, P.prefer_waves);
    if (st != TCHGEO_OK) return st;
  }
  // ---- 3. per relation: edges in generation order ----------------------------------------------------
  for (int r = 0; r < R && G > 0; ++r) {
    const int s = a->rel_src[r], t = a->rel_dst[r];
    if (P.S[s] == 0 || a->num_neg == 0) continue;
    TCHGEO_REQUIRE(a->rows[r] && a->cols[r], "rows/cols[%d] is NULL", r);
    const NegFlag flag{cand, crel, d_rel_dst, r, 0};
    size_t cub = P.cub_bytes;
    TCHGEO_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(ws + P.off_cub, cub, flag_iterator(flag), ranks, G, stream));
    neg_scatter_rel_kernel<<<ggrid, NEG_THREADS, 0, stream>>>(flag, ranks, tpos, P.slot_base[s],
                                                             P.slot_base[s] + P.S[s] * a->num_neg, G, a->num_neg,
                                                             local + P.seq_off[t] + P.S[t], a->rows[r], a->cols[r],
                                                             d_elen + r);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  std::vector<int64_t> h((size_t)(3 * T + R));
  uint32_t e = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_acc, h.size() * 8, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&e, d_err, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  for (int t = 0; t < T; ++t) a->samples_len[t] = h[(size_t)(2 * T + t)];
  for (int r = 0; r < R; ++r) a->edges_len[r] = h[(size_t)(3 * T + r)];
  return status_from_dev_err(e);
}
