// Multi-hop fixed-fanout neighbor sampling for sm_100a (B200).
//
// Replaces src/algo/neighbor_sampling.rs:162-230 (homogenous) and :233-356 (heterogenous) of the
// reference, together with the samplers of src/utils/sampling.rs:6-69, behind the C ABI declared in
// include/tchgeo_cuda.h.  Output layout is the reference's computation tree (no dedup): every
// sampled edge appends one node, rows[e] = index of that node in samples[src], cols[e] = index of the
// frontier node in samples[dst], edge_index[e] = CSC position (neighbor_sampling.rs:210-218).
//
// One kernel launch per (hop, relation) processes that hop's frontier of *all* batches:
//   tile = up to 128 frontier nodes of one batch; a 128-thread CTA works through two tiles
//   A. coalesced load of frontier ids, gather of the (ptrs[w], ptrs[w+1]) pair (both tiles up front), per-node
//      count cnt = min(deg, k) (or k*[deg>0] with replacement), prefix sums by warp scans
//   B. the tile's exclusive output offset inside its batch comes from a decoupled look-back over the
//      batch's earlier tiles (single pass, no separate count/scan kernels, no host sync)
//   C. sampling decisions in shared memory, one word per output slot that ends up holding the position of the
//      chosen neighbour inside the neighbourhood
//        UNIFORM : the serial reservoir of sampling.rs:6-26 is replayed exactly in parallel:
//                  slot s ends up holding the item of the LAST step i>=k whose draw j_i == s, else
//                  item s.  Steps are independent given i, so the tile's (node, 8-step chunk) pairs
//                  are flattened over the CTA; a chunk is Philox blocks (seed; pos, 2c | 2c+1, batch) and
//                  red.shared.max(slot, step) for the hits (the slot starts as s, and steps >= k > s).
//        REPLACE : k iid draws per non-empty neighbourhood (sampling.rs:57-69), one Philox block per four.
//        WEIGHTED: step i fires iff u*w_sum_i < w_i and then overwrites a uniform slot (sampling.rs:28-55);
//                  last firing writer per slot wins via atomicMax.  w_sum_i comes from serial per-column
//                  prefix sums (eight lanes per 8-step chunk), or, without them, from a warp scan per node.
//   D. thread-per-output-edge epilogue: one random 4/8-byte gather from row_indices and four fully
//      coalesced 8-byte streaming stores (samples, rows, cols, edge_index).
// HBM-bound integer work: no tensor cores.  See DESIGN.md for the byte model, the RNG contract and what was
// measured and rejected (warp tiles, four tiles per CTA, deferred stores).
#include <cub/block/block_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "graph.cuh"

namespace tchgeo {
namespace {

constexpr int HOP_THREADS = 128;          // threads (= max frontier nodes) per tile
constexpr int HOP_TPC = 2;                // tiles per CTA: the colptr gathers of both are requested up front
constexpr int HOP_DEFAULT_MIN_BLOCKS = 10;  // CTAs per SM the register budget is set for (48 registers)
constexpr int HOP_U = 5;                  // gathers in flight per thread in phase D
constexpr int MAX_TILE_EDGES = 8192;   // shared-memory slots per tile when fanout <= 8192
constexpr int MAX_FANOUT = 32768;      // one node per tile above 8192; bounded by shared memory
constexpr uint64_t ST_FLAG_AGG = 1ull << 62;
constexpr uint64_t ST_FLAG_INCL = 2ull << 62;
constexpr uint64_t ST_VAL_MASK = (1ull << 62) - 1;
constexpr int LIGHT_CHUNKS_MAX = 16;   // nodes needing more 8-step draw chunks are strided by a warp each

struct HopParams {
  const int64_t* ptrs;
  const int64_t* indices;
  const int32_t* indices32;  // optional compressed replica of `indices`
  const double* weights;
  const double2* wrec;       // optional (weight, serial per-column prefix sum) records (graph handle): one 16-byte load per step
  int64_t num_cols;
  int64_t nnz;               // entries of indices / weights / timestamps (INT64_MAX = unchecked)
  const int64_t* dst_samples;  // frontier ids live here (may alias src_samples)
  int64_t dst_stride;
  int64_t* src_samples;        // sampled nodes are appended here
  int64_t src_stride;
  int64_t* rows;
  int64_t* cols;
  int64_t* eidx;
  int64_t e_stride;
  const int64_t* fr_begin;     // [B] frontier window in dst_samples
  const int64_t* fr_end;       // [B]
  const int64_t* src_len_in;   // [B] len(samples[src]) before this launch
  int64_t* src_len_out;        // [B] ... after
  const int64_t* e_len_in;     // [B] len(edges[rel]) before this launch
  int64_t* e_len_out;          // [B] ... after
  uint64_t* status;            // [B * tiles_per_batch] look-back words, zero-initialised
  uint32_t* ticket;            // zero-initialised tile dispenser
  uint32_t* err;
  int32_t num_batches;
  int32_t tiles_per_batch;
  int32_t tile_nodes;
  int32_t tile_edges;          // tile_nodes * fanout
  int32_t fanout;
  uint32_t key0, key1;
  uint32_t rk[20];              // Philox round keys (key0 + r*W0, key1 + r*W1): constant-bank operands, no registers
  uint32_t rel;
  uint32_t batch_base;
  uint32_t total_tiles;         // num_batches * tiles_per_batch
  // temporal filter (src/algo/neighbor_sampling.rs:36-77); filter_mode 0 = none
  int32_t filter_mode;          // 1 static, 2 relative, 3 dynamic
  int32_t filter_forward;
  int64_t win_lo, win_hi;
  const int64_t* timestamps;    // [nnz] per CSC position
  const int64_t* dst_states;    // [B, dst_stride] state of every sample of the dst type
  int64_t* src_states;          // [B, src_stride] states of appended samples are written here
};

// err[0] |= DEV_ERR_INDEX when an id does not fit; WITH_MAX: err[1] = max id (the relabel stage's id bound)
template <bool WITH_MAX>
__global__ void __launch_bounds__(256) compress_kernel(const int64_t* __restrict__ src, int64_t n,
                                                       int32_t* __restrict__ dst, uint32_t* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  uint32_t mx = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t v = src[i];
    bad |= (v < 0 || v > 0x7fffffffll);
    dst[i] = (int32_t)v;
    if (WITH_MAX && (uint32_t)v > mx) mx = (uint32_t)v;
  }
  if (bad) atomicOr(err, DEV_ERR_INDEX);
  if (WITH_MAX) {
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(err + 1, mx);
  }
}

__global__ void fill_i64_kernel(int64_t* p, int64_t v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// per-node record kept in shared memory (one 8-byte load in the hot loops)
struct __align__(8) NodeRec {
  uint32_t deg;    // neighbourhood size
  uint16_t off;    // exclusive output offset inside the tile
  uint16_t choff;  // exclusive draw-block offset (light nodes)
};

struct TileHdr {  // written by thread 0, read by everyone after the first barrier
  int64_t fb, F, e_in, s_in;
  int b, t;
};

__device__ __forceinline__ uint32_t smem_atom_add(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;"
               : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}

// steps step0 .. step0+3 of the serial reservoir (sampling.rs:17-23) for one node: draw j uniform in
// [0, step); a hit (j < k) overwrites slot j; the LAST hit of a slot wins -> atomicMax on the step index.
__device__ __forceinline__ void reservoir_block(const Philox4& r, uint32_t step0, uint32_t deg, uint32_t k,
                                                uint32_t* slots) {
#pragma unroll
  for (uint32_t u = 0; u < 4; ++u) {
    const uint32_t step = step0 + u;
    const uint32_t j = __umulhi(pick4(r, u), step);
    const bool hit = (step < deg) & (j < k);
    if (hit) atomicMax(slots + j, step);
  }
}

// Same as reservoir_block, but straight-line: `slots_sa` is a 32-bit shared-state-space address and the
// update is a predicated red.shared.max (no branch, no generic->shared address conversion per hit).
__device__ __forceinline__ void reservoir_block_sa(const Philox4& r, uint32_t step0, uint32_t deg, uint32_t k,
                                                   uint32_t slots_sa) {
#pragma unroll
  for (uint32_t u = 0; u < 4; ++u) {
    const uint32_t step = step0 + u;
    const uint32_t j = __umulhi(pick4(r, u), step);
    const uint32_t hit = (uint32_t)((step < deg) & (j < k));
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.max.u32 [%0], %1;\n\t}"
        ::"r"(slots_sa + 4u * j), "r"(step), "r"(hit)
        : "memory");
  }
}

// Philox4x32-10 with the round keys taken from the kernel parameters (constant bank): same function as
// philox4x32_10(..., p.key0, p.key1), but the 18 bumped keys do not occupy registers for the whole draw loop.
__device__ __forceinline__ Philox4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const HopParams& p) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ p.rk[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ p.rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}

template <int KIND, int MINB, bool I32>
__global__ void __launch_bounds__(HOP_THREADS, MINB) hop_kernel(const HopParams p) {
  constexpr int NT = HOP_THREADS, TPC = HOP_TPC, NW = NT / 32;
  __shared__ int64_t s_start[NT];
  __shared__ NodeRec s_rec[NT];
  __shared__ uint8_t s_chown[LIGHT_CHUNKS_MAX * NT];  // 8-step draw chunk -> owning node
  __shared__ uint8_t s_heavy[NT];
  __shared__ TileHdr s_hdr[TPC];
  __shared__ int64_t s_pre_start[TPC][NT];  // prefetched colptr data of the CTA's TPC tiles
  __shared__ uint32_t s_pre_deg[TPC][NT];
  __shared__ __align__(16) uint32_t s_wtot[NW];  // per-warp totals of the tile's prefix sums
  __shared__ uint32_t s_work;     // dynamic work counter of the draw phase
  __shared__ uint32_t s_nheavy;
  __shared__ int64_t s_excl;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(dyn_smem);
  uint8_t* s_owner = reinterpret_cast<uint8_t*>(s_slot + p.tile_edges);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int TN = p.tile_nodes;
  const uint32_t k = (uint32_t)p.fanout;
  if (tid == 0) {
    // Tiles are handed out in start order and TILE-MAJOR (ticket -> tile t of batch b = ticket % B):
    // the tiles in flight at any moment belong to different batches, so the per-batch look-back
    // chains advance independently, and every tile this one waits on (same batch, smaller t) holds a
    // smaller ticket, i.e. belongs to a CTA that is already running (or to this CTA, earlier in its loop).
    const uint32_t base = atomicAdd(p.ticket, (uint32_t)TPC);
#pragma unroll
    for (int i = 0; i < TPC; ++i) {
      const uint32_t ticket = base + i;
      TileHdr h;
      h.fb = 0; h.F = 0; h.e_in = 0; h.s_in = 0; h.b = 0; h.t = 1;  // t=1, F=0: an empty tile that is not "last"
      if (ticket < p.total_tiles) {
        h.t = (int)(ticket / (uint32_t)p.num_batches);
        h.b = (int)(ticket - (uint32_t)h.t * (uint32_t)p.num_batches);
        h.fb = p.fr_begin[h.b];
        int64_t fe = p.fr_end[h.b];
        if (fe > p.dst_stride) fe = p.dst_stride;  // only after a capacity error upstream
        h.F = fe > h.fb ? fe - h.fb : 0;
        h.e_in = p.e_len_in[h.b];
        h.s_in = p.src_len_in[h.b];
      }
      s_hdr[i] = h;
    }
    s_work = 0u;
    s_nheavy = 0u;
  }
  __syncthreads();

  // ---- A: frontier ids and colptr pairs of ALL the CTA's tiles are requested up front, so the two
  //         dependent gathers of tile i+1 overlap the processing of tile i --------------------------
  {
    int64_t w[TPC], st[TPC], en[TPC];
#pragma unroll
    for (int i = 0; i < TPC; ++i) {
      const int64_t node0 = (int64_t)s_hdr[i].t * TN;
      w[i] = -2;
      if (node0 + tid < s_hdr[i].F && tid < TN)
        w[i] = p.dst_samples[(int64_t)s_hdr[i].b * p.dst_stride + s_hdr[i].fb + node0 + tid];
    }
    const uint64_t keep = l2_policy_evict_last();  // colptr (8 B/node) should live in the 126 MB L2
#pragma unroll
    for (int i = 0; i < TPC; ++i) {
      st[i] = 0; en[i] = 0;
      if (w[i] >= 0 && w[i] < p.num_cols) {
        st[i] = ld_gather64_keep_i64(p.ptrs + w[i], keep);
        en[i] = ld_gather64_keep_i64(p.ptrs + w[i] + 1, keep);
      } else if (w[i] != -2) {
        atomicOr(p.err, DEV_ERR_INDEX);  // reference: slice index panic (quirk Q10)
      }
    }
#pragma unroll
    for (int i = 0; i < TPC; ++i) {
      const int64_t d = en[i] - st[i];
      uint32_t deg = 0;
      // a column that ends beyond the edge arrays: the reference panics on the slice (graph.rs:72-78)
      if (d < 0 || d > 0x7fffffffll || st[i] < 0 || en[i] > p.nnz) atomicOr(p.err, DEV_ERR_INDEX);
      else deg = (uint32_t)d;
      s_pre_start[i][tid] = st[i];
      s_pre_deg[i][tid] = deg;
    }
  }

  using val_t = typename std::conditional<I32, int32_t, int64_t>::type;

#pragma unroll 1
  for (int ti = 0; ti < TPC; ++ti) {
  if (ti) __syncthreads();  // the shared tables of the previous tile are free again
  const int b = s_hdr[ti].b, t = s_hdr[ti].t;
  const uint32_t F = (uint32_t)s_hdr[ti].F;       // < 2^32: samples_stride is
  const uint32_t node0 = (uint32_t)t * (uint32_t)TN;
  const int nn = F > node0 ? (int)min((uint32_t)TN, F - node0) : 0;
  const bool is_last = (nn > 0 && node0 + (uint32_t)nn == F) || (F == 0 && t == 0);
  if (nn == 0 && !is_last) continue;

  uint32_t cnt = 0;
  uint32_t deg = 0;
  uint32_t nch = 0;  // 8-step draw chunks this node needs (UNIFORM, WEIGHTED on prefix sums)
  bool heavy = false;
  if (tid < nn) {
    deg = s_pre_deg[ti][tid];
    if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
      cnt = deg > 0 ? k : 0;  // exactly k picks, even when deg < k (quirk Q3)
    } else {
      cnt = deg < k ? deg : k;
      if (k == 0 && deg > 0) atomicOr(p.err, DEV_ERR_PANIC);  // gen_range(0..0), sampling.rs:19
      // UNIFORM and WEIGHTED-with-prefix-sums evaluate steps k .. deg-1 independently: 8-step work items
      if ((KIND == TCHGEO_SAMPLER_UNIFORM || (KIND == TCHGEO_SAMPLER_WEIGHTED && p.wrec)) && deg > k) {
        nch = (deg - k + 7u) >> 3;
        heavy = nch > (uint32_t)LIGHT_CHUNKS_MAX;
      }
    }
  }
  // one 32-bit scan carries both prefix sums: low 16 bits = outputs (<= 32768 per tile),
  // high 16 bits = draw chunks of the light nodes (<= LIGHT_CHUNKS_MAX * NT).  Warp scans + the NW warp totals.
  const uint32_t light_chunks = heavy ? 0u : nch;
  const uint32_t sv = cnt | (light_chunks << 16);
  uint32_t incl = sv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t pexcl = incl - sv, ptotal = 0;
  {
    static_assert(NW == 4, "the warp totals are read as one 16-byte vector");
    const uint4 wt = *reinterpret_cast<const uint4*>(s_wtot);
    pexcl += (warp > 0 ? wt.x : 0u) + (warp > 1 ? wt.y : 0u) + (warp > 2 ? wt.z : 0u);
    ptotal = wt.x + wt.y + wt.z + wt.w;
  }
  const uint32_t off = pexcl & 0xffffu;
  const uint32_t total = ptotal & 0xffffu;
  const uint32_t choff = pexcl >> 16;
  const uint32_t Q = ptotal >> 16;

  // publish this tile's aggregate as early as possible
  uint64_t* my_status = p.status + (size_t)b * p.tiles_per_batch;
  if (tid == 0) st_relaxed_u64(my_status + t, (t == 0 ? ST_FLAG_INCL : ST_FLAG_AGG) | (uint64_t)total);

  // ---- C1: per-node set-up of the shared tables (own node only: needs no barrier) ---------------
  s_start[tid] = s_pre_start[ti][tid];
  s_rec[tid] = NodeRec{deg, (uint16_t)off, (uint16_t)choff};
  if (tid < nn) {
    // slot s starts as "item s" and is raised to the step index of every hit (steps >= k > s), so after the
    // draws it holds the position of the chosen neighbour directly: last hit, else item s (sampling.rs:17-23)
    for (uint32_t s = 0; s < cnt; ++s) {
      s_owner[off + s] = (uint8_t)tid;
      if (KIND != TCHGEO_SAMPLER_UNIFORM_REPLACE) s_slot[off + s] = s;
    }
    if (KIND == TCHGEO_SAMPLER_UNIFORM || KIND == TCHGEO_SAMPLER_WEIGHTED) {  // (no work items without prefix sums)
      for (uint32_t c = 0; c < light_chunks; ++c) s_chown[choff + c] = (uint8_t)tid;
      if (heavy) s_heavy[atomicAdd(&s_nheavy, 1u)] = (uint8_t)tid;
    }
  }
  __syncthreads();

  // ---- B: decoupled look-back over this batch's earlier tiles (warp 0), overlapped with C2 ------
  if (tid < 32) {
    int64_t excl = 0;
    if (t > 0) {
      int j = t - 1;
      uint32_t spins = 0;
      while (true) {
        const int idx = j - lane;
        const uint64_t v = idx >= 0 ? ld_relaxed_u64(my_status + idx) : ST_FLAG_INCL;
        const uint32_t flag = (uint32_t)(v >> 62);
        const uint32_t incl_mask = __ballot_sync(0xffffffffu, flag == 2u);
        const uint32_t inval_mask = __ballot_sync(0xffffffffu, flag == 0u);
        const int first_incl = incl_mask ? __ffs(incl_mask) - 1 : 32;
        const int first_inval = inval_mask ? __ffs(inval_mask) - 1 : 32;
        if (first_inval < first_incl) {  // a predecessor we need has not published yet
          if (++spins > (1u << 24)) {
            if (lane == 0) atomicOr(p.err, DEV_ERR_WATCHDOG);
            break;
          }
          __nanosleep(32);
          continue;
        }
        int64_t val = lane <= first_incl ? (int64_t)(v & ST_VAL_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (first_incl < 32) break;
        j -= 32;
      }
      if (lane == 0) st_relaxed_u64(my_status + t, ST_FLAG_INCL | (uint64_t)(excl + total));
    }
    if (lane == 0) s_excl = excl;
  }

  // ---- C2: sampling decisions in shared memory --------------------------------------------------
  const uint32_t pos0 = (uint32_t)s_hdr[ti].fb + node0;
  const uint32_t batch = p.batch_base + (uint32_t)b;

  if (KIND == TCHGEO_SAMPLER_UNIFORM) {
    const uint32_t tag = TAG_RESERVOIR | (p.rel << 8);
    const uint32_t slot_sa = (uint32_t)__cvta_generic_to_shared(s_slot);
    // Light nodes: (node, 8-step chunk) work items, grabbed 128 at a time so that warp 0 joins in after its
    // look-back.  Chunk c of node n is Philox blocks 2c and 2c+1 of the contract: steps k+8c .. k+8c+7 of the
    // serial reservoir; the second block is skipped when it lies past the neighbourhood.
    while (true) {
      uint32_t base = 0;
      if (lane == 0) base = smem_atom_add(&s_work, 128u);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= Q) break;
      const uint32_t qend = min(base + 128u, Q);
#pragma unroll 1
      for (uint32_t q = base + lane; q < qend; q += 32) {
        const uint32_t n = s_chown[q];
        const NodeRec rec = s_rec[n];
        const uint32_t c = q - rec.choff;
        const uint32_t step0 = k + 8u * c;
        const uint32_t sa = slot_sa + 4u * rec.off;
        // (the second block stays a predicated tail: an if/else with both blocks interleaved in one branch makes
        //  every mixed warp run three Philox bodies instead of two: hop 3 1.65 -> 1.73 ms)
        reservoir_block_sa(philox_rk(pos0 + n, 2u * c, batch, tag, p), step0, rec.deg, k, sa);
        if (step0 + 4u < rec.deg)
          reservoir_block_sa(philox_rk(pos0 + n, 2u * c + 1u, batch, tag, p), step0 + 4u, rec.deg, k, sa);
      }
    }
    // Heavy nodes (deg > k + 8*LIGHT_CHUNKS_MAX): the whole CTA strides over one node's 4-step blocks, so that no
    // warp is left alone with a hub while the others wait at the barrier below.
    const uint32_t nheavy = s_nheavy;
    for (uint32_t h = 0; h < nheavy; ++h) {
      const uint32_t n = s_heavy[h];
      const NodeRec rec = s_rec[n];
      const uint32_t nb = (rec.deg - k + 3u) >> 2;
#pragma unroll 1
      for (uint32_t c = tid; c < nb; c += NT) {
        const Philox4 r = philox_rk(pos0 + n, c, batch, tag, p);
        reservoir_block_sa(r, k + 4u * c, rec.deg, k, slot_sa + 4u * rec.off);
      }
    }
  } else if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
    // k iid picks per non-empty neighbourhood (sampling.rs:57-69): work item = (node, Philox block c) -> slots
    // 4c .. 4c+3, one Philox call per four picks
    const uint32_t rtag = TAG_REPLACE | (p.rel << 8);
    const uint32_t bpn = (k + 3u) >> 2;  // Philox blocks per node
#pragma unroll 1
    for (uint32_t q = tid; q < (uint32_t)nn * bpn; q += NT) {
      const uint32_t n = q / bpn, c = q - n * bpn;
      const NodeRec rec = s_rec[n];
      if (rec.deg == 0) continue;
      const Philox4 r = philox_rk(pos0 + n, c, batch, rtag, p);
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u)
        if (4u * c + u < k) s_slot[rec.off + 4u * c + u] = __umulhi(pick4(r, u), rec.deg);  // sampling.rs:64
    }
  } else if (KIND == TCHGEO_SAMPLER_WEIGHTED) {
    const uint32_t wtag = TAG_WEIGHTED | (p.rel << 8);
    if (p.wrec) {
      // w_sum at every step comes from the precomputed serial prefix sums: no scan, the comparison below sees
      // exactly the reference's f64 values whatever the weights are (sampling.rs:47-52), and the steps are
      // independent, so they are spread over the CTA as 8-step work items exactly like the uniform reservoir.
      auto weighted_step = [&](uint32_t n, const NodeRec& rec, uint32_t item) {
        const double2 wr = __ldg(p.wrec + s_start[n] + item);
        const double w = wr.x, w_sum = wr.y;
        if (!(w_sum > 0.0)) {
          atomicOr(p.err, DEV_ERR_PANIC);  // gen_range(0.0..w_sum) on an empty range
          return;
        }
        const Philox4 r = philox_rk(pos0 + n, item, batch, wtag, p);
        const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
        const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
        if (__dmul_rn(u, w_sum) < w) atomicMax(s_slot + rec.off + __umulhi(r.z, k), item);  // :49-52
      };
      // 8 lanes per chunk, one step per lane: the eight (weight, prefix sum) records of a chunk are 128 contiguous
      // bytes, one line.  Four chunks per lane group are requested before the first one is used (memory-level parallelism).
      while (true) {
        uint32_t base = 0;
        if (lane == 0) base = smem_atom_add(&s_work, 128u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= Q) break;
        const uint32_t qend = min(base + 128u, Q);
#pragma unroll 1
        for (uint32_t q0 = base + (lane >> 3); q0 < qend; q0 += 16) {
          uint32_t nn4[4], item4[4];
          double w4[4], ws4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t q = q0 + 4u * i;
            nn4[i] = 0xffffffffu;
            if (q < qend) {
              const uint32_t n = s_chown[q];
              const NodeRec rec = s_rec[n];
              const uint32_t item = k + 8u * (q - rec.choff) + (lane & 7u);
              if (item < rec.deg) {
                nn4[i] = n;
                item4[i] = item;
                const double2 wr = __ldg(p.wrec + s_start[n] + item);
                w4[i] = wr.x;
                ws4[i] = wr.y;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (nn4[i] == 0xffffffffu) continue;
            if (!(ws4[i] > 0.0)) {
              atomicOr(p.err, DEV_ERR_PANIC);  // gen_range(0.0..w_sum) on an empty range
              continue;
            }
            const Philox4 r = philox_rk(pos0 + nn4[i], item4[i], batch, wtag, p);
            const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
            const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
            if (__dmul_rn(u, ws4[i]) < w4[i])
              atomicMax(s_slot + s_rec[nn4[i]].off + __umulhi(r.z, k), item4[i]);  // :49-52
          }
        }
      }
      const uint32_t nheavy = s_nheavy;
      for (uint32_t h = 0; h < nheavy; ++h) {
        const uint32_t n = s_heavy[h];
        const NodeRec rec = s_rec[n];
        for (uint32_t item = k + tid; item < rec.deg; item += NT) weighted_step(n, rec, item);
      }
    } else {
    for (int n = warp; n < nn; n += NW) {
      const NodeRec rec = s_rec[n];
      const uint32_t dn = rec.deg;
      if (dn <= k) continue;
      const double* wp = p.weights + s_start[n];
      uint32_t* slots = s_slot + rec.off;
      double carry = 0.0;
      for (uint32_t base = 0; base < dn; base += 32) {
        const uint32_t item = base + lane;
        const double w = item < dn ? __ldg(wp + item) : 0.0;
        double incl_w = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, incl_w, o);
          if (lane >= o) incl_w += up;
        }
        const double w_sum = carry + incl_w;  // sampling.rs:48
        if (item >= k && item < dn) {
          if (!(w_sum > 0.0)) {
            atomicOr(p.err, DEV_ERR_PANIC);  // gen_range(0.0..w_sum) on an empty range
          } else {
            const Philox4 r = philox4x32_10(pos0 + (uint32_t)n, item, batch, wtag, p.key0, p.key1);
            const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
            const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
            if (__dmul_rn(u, w_sum) < w) atomicMax(slots + __umulhi(r.z, k), item);  // :49-52
          }
        }
        carry = __shfl_sync(0xffffffffu, w_sum, 31);
      }
    }
    }
  }
  __syncthreads();

  const int64_t excl = s_excl;
  const int64_t e_base = s_hdr[ti].e_in + excl;
  const int64_t s_base = s_hdr[ti].s_in + excl;
  if (tid == 0) {  // reset for the CTA's next tile (ordered by the barrier at the top of the loop)
    s_work = 0u;
    s_nheavy = 0u;
  }
  if (is_last && tid == 0) {
    p.e_len_out[b] = e_base + total;
    p.src_len_out[b] = s_base + total;
  }
  if (e_base + total > p.e_stride || s_base + total > p.src_stride) {
    if (tid == 0) atomicOr(p.err, DEV_ERR_CAPACITY);
    continue;
  }
  if (total == 0) continue;

  // ---- D: one thread per output edge; HOP_U gathers in flight per thread (the whole tile when fanout <= HOP_U).
  //         The three stores that do not depend on the gather are issued while it is in flight.  cols and rows
  //         are < 2^32 (samples_stride is), so their values are formed in 32 bits. -------------------------------
  int64_t* const pe = p.eidx + ((int64_t)b * p.e_stride + e_base);
  int64_t* const pc = p.cols + ((int64_t)b * p.e_stride + e_base);
  int64_t* const pr = p.rows + ((int64_t)b * p.e_stride + e_base);
  int64_t* const ps = p.src_samples + ((int64_t)b * p.src_stride + s_base);
  const uint32_t col0 = (uint32_t)s_hdr[ti].fb + node0;
  const uint32_t row0 = (uint32_t)s_base;
#pragma unroll 1
  for (uint32_t e0 = tid; e0 < total; e0 += NT * HOP_U) {
    val_t val[HOP_U];
    if (e0 - lane + 31u + (HOP_U - 1) * NT < total) {  // all HOP_U edges of every lane of this warp exist: no
                                                       // predicates, and no warp runs both branches
      uint32_t own[HOP_U];
      int64_t ptr[HOP_U];
#pragma unroll
      for (uint32_t u = 0; u < HOP_U; ++u) {
        const uint32_t e = e0 + u * NT;
        own[u] = s_owner[e];
        ptr[u] = s_start[own[u]] + s_slot[e];
        if (I32) val[u] = ld_gather64_i32(p.indices32 + ptr[u]);
        else val[u] = ld_gather64_i64(p.indices + ptr[u]);
      }
#pragma unroll
      for (uint32_t u = 0; u < HOP_U; ++u) {
        const uint32_t e = e0 + u * NT;
        st_cs_i64(pe + e, ptr[u]);
        st_cs_i64(pc + e, (int64_t)(uint64_t)(col0 + own[u]));
        st_cs_i64(pr + e, (int64_t)(uint64_t)(row0 + e));
      }
    } else {
      uint32_t rel[HOP_U], own[HOP_U];
#pragma unroll
      for (uint32_t u = 0; u < HOP_U; ++u) {
        const uint32_t e = min(e0 + u * NT, total - 1u);  // clamped: tail lanes redo the last edge
        own[u] = s_owner[e];
        rel[u] = s_slot[e];
        const int64_t ptr = s_start[own[u]] + rel[u];
        val[u] = 0;
        if (e0 + u * NT < total) {
          if (I32) val[u] = ld_gather64_i32(p.indices32 + ptr);
          else val[u] = ld_gather64_i64(p.indices + ptr);
        }
      }
#pragma unroll
      for (uint32_t u = 0; u < HOP_U; ++u) {
        const uint32_t e = e0 + u * NT;
        if (e < total) {
          st_cs_i64(pe + e, s_start[own[u]] + rel[u]);
          st_cs_i64(pc + e, (int64_t)(uint64_t)(col0 + own[u]));
          st_cs_i64(pr + e, (int64_t)(uint64_t)(row0 + e));
        }
      }
    }
#pragma unroll
    for (uint32_t u = 0; u < HOP_U; ++u)
      if (e0 + u * NT < total) st_cs_i64(ps + e0 + u * NT, (int64_t)val[u]);  // read again only by the next hop
  }
  }  // tiles of this CTA
}

// ---------------------------------------------------------------------------------------------
// Temporal-filter variant (src/algo/neighbor_sampling.rs:36-77 + :204-217).
// The samplers see only the edges that pass the filter, in CSC order, so every position below is a
// position AMONG THE PASSING EDGES.  One warp per frontier node: pass 1 counts the passing edges,
// the tile then takes its output offset from the same look-back protocol as hop_kernel, pass 2 makes the
// sampling decisions (same Philox counters, steps indexed by passing position) and a final sweep over
// the neighbourhood emits the chosen edges and the new per-sample states.
// ---------------------------------------------------------------------------------------------
constexpr int FT_THREADS = 128;
constexpr int FT_MASKW = 2;   // passing-edge masks of a node's first 64 edges are kept in shared memory by pass 1
constexpr uint32_t FT_LIGHT = 32u * FT_MASKW;  // nodes up to this degree are handled by ONE THREAD each, hubs by a warp

// STATIC: a fixed window on the timestamp (:59); RELATIVE / DYNAMIC: a window on the distance to the sample's state (:60-65)
template <bool STATIC>
__device__ __forceinline__ bool filter_pass(const HopParams& p, int64_t t, int64_t state) {
  if (STATIC) return p.win_lo <= t && t <= p.win_hi;
  int64_t d = t - state;
  if (!p.filter_forward) d = -d;
  return p.win_lo <= d && d <= p.win_hi;
}

template <int KIND, bool STATIC>
__global__ void __launch_bounds__(FT_THREADS) hop_filtered_kernel(const HopParams p) {
  using BlockScan = cub::BlockScan<uint32_t, FT_THREADS>;
  __shared__ typename BlockScan::TempStorage scan_tmp;
  __shared__ int64_t s_start[FT_THREADS], s_state[FT_THREADS];
  __shared__ uint32_t s_deg[FT_THREADS], s_pass[FT_THREADS], s_off[FT_THREADS];
  __shared__ TileHdr s_hdr;
  __shared__ int64_t s_excl;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(dyn_smem);  // [tile_edges]
  uint32_t* s_mask = s_slot + p.tile_edges;                  // [FT_THREADS * FT_MASKW] passing-edge masks from pass 1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = FT_THREADS / 32;
  const int TN = p.tile_nodes;
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(p.ticket, 1u);
    const int t = (int)(ticket / (uint32_t)p.num_batches);
    const int b = (int)(ticket - (uint32_t)t * (uint32_t)p.num_batches);
    const int64_t fb = p.fr_begin[b];
    int64_t fe = p.fr_end[b];
    if (fe > p.dst_stride) fe = p.dst_stride;
    s_hdr.fb = fb; s_hdr.F = fe > fb ? fe - fb : 0;
    s_hdr.e_in = p.e_len_in[b]; s_hdr.s_in = p.src_len_in[b];
    s_hdr.b = b; s_hdr.t = t;
  }
  __syncthreads();
  const int b = s_hdr.b, t = s_hdr.t;
  const int64_t fb = s_hdr.fb, F = s_hdr.F;
  const uint32_t k = (uint32_t)p.fanout;
  const int64_t node0 = (int64_t)t * TN;
  const int nn = (int)max((int64_t)0, min((int64_t)TN, F - node0));
  const bool is_last = (nn > 0 && node0 + nn == F) || (F == 0 && t == 0);
  if (nn == 0 && !is_last) return;

  // ---- pass 1: number of passing edges per node ------------------------------------------------
  // Light nodes (deg <= 64: nearly all of them, the mean degree is 25) take ONE THREAD each: the thread walks its
  // column's timestamps and keeps the passing mask in two registers.  With a warp per node the kernel was bound by the
  // fixed per-node cost of the warp loop (id -> colptr -> timestamps chain, 71 % issue utilisation, most lanes idle);
  // a thread per node shares that cost over 32 nodes.  Hubs keep the warp loop below.
  __shared__ uint8_t s_hub[FT_THREADS];
  __shared__ uint32_t s_nhub;
  if (tid == 0) s_nhub = 0u;
  __syncthreads();
  if (tid < nn) {
    const int64_t gi = (int64_t)b * p.dst_stride + fb + node0 + tid;
    const int64_t w = p.dst_samples[gi];
    const int64_t state = p.dst_states[gi];
    int64_t start = 0;
    uint32_t deg = 0;
    if (w < 0 || w >= p.num_cols) {
      atomicOr(p.err, DEV_ERR_INDEX);
    } else {
      start = __ldg(p.ptrs + w);
      const int64_t d = __ldg(p.ptrs + w + 1) - start;
      if (d < 0 || d > 0x7fffffffll || start < 0 || start + d > p.nnz) atomicOr(p.err, DEV_ERR_INDEX);
      else deg = (uint32_t)d;
    }
    s_start[tid] = start; s_state[tid] = state; s_deg[tid] = deg;
    if (deg <= FT_LIGHT) {
      uint32_t lo = 0, hi = 0;
      const int64_t* ts = p.timestamps + start;
      // eight timestamps are requested before the first is compared: the loop is one DRAM latency per batch and thread
      // (29 % of the kernel's stall samples sat on the compare of the one-by-one version, r2_hop_filtered.ncu-rep)
      for (uint32_t e0 = 0; e0 < deg; e0 += 8u) {
        int64_t tt[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) tt[u] = e0 + u < deg ? __ldg(ts + e0 + u) : 0;
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) {
          const uint32_t e = e0 + u;
          if (e >= deg) break;
          const uint32_t ok = filter_pass<STATIC>(p, tt[u], state) ? 1u : 0u;
          if (e < 32u) lo |= ok << e; else hi |= ok << (e - 32u);
        }
      }
      s_mask[tid * FT_MASKW] = lo;
      s_mask[tid * FT_MASKW + 1] = hi;
      s_pass[tid] = __popc(lo) + __popc(hi);
    } else {
      s_hub[atomicAdd(&s_nhub, 1u)] = (uint8_t)tid;
    }
  }
  __syncthreads();
  for (uint32_t hb = warp; hb < s_nhub; hb += NWARP) {   // hubs: one warp sweeps the column
    const int n = s_hub[hb];
    const int64_t start = s_start[n], state = s_state[n];
    const uint32_t deg = s_deg[n];
    uint32_t npass = 0;
    for (uint32_t base = 0; base < deg; base += 32) {
      const uint32_t item = base + lane;
      const bool ok = item < deg && filter_pass<STATIC>(p, __ldg(p.timestamps + start + item), state);
      const uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0 && base < FT_LIGHT) s_mask[n * FT_MASKW + (base >> 5)] = m;  // the select pass reuses it
      npass += __popc(m);
    }
    if (lane == 0) s_pass[n] = npass;
  }
  __syncthreads();
  uint32_t cnt = 0;
  if (tid < nn) {
    const uint32_t np = s_pass[tid];
    if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) cnt = np > 0 ? k : 0;
    else {
      cnt = np < k ? np : k;
      if (k == 0 && np > 0) atomicOr(p.err, DEV_ERR_PANIC);
    }
  }
  uint32_t off, total;
  BlockScan(scan_tmp).ExclusiveSum(cnt, off, total);
  s_off[tid] = off;

  // ---- look-back (same protocol as hop_kernel) --------------------------------------------------
  uint64_t* my_status = p.status + (size_t)b * p.tiles_per_batch;
  if (tid < 32) {
    int64_t excl = 0;
    if (t == 0) {
      if (lane == 0) st_relaxed_u64(my_status, ST_FLAG_INCL | (uint64_t)total);
    } else {
      if (lane == 0) st_relaxed_u64(my_status + t, ST_FLAG_AGG | (uint64_t)total);
      int j = t - 1;
      uint32_t spins = 0;
      while (true) {
        const int idx = j - lane;
        const uint64_t v = idx >= 0 ? ld_relaxed_u64(my_status + idx) : ST_FLAG_INCL;
        const uint32_t flag = (uint32_t)(v >> 62);
        const uint32_t incl_mask = __ballot_sync(0xffffffffu, flag == 2u);
        const uint32_t inval_mask = __ballot_sync(0xffffffffu, flag == 0u);
        const int first_incl = incl_mask ? __ffs(incl_mask) - 1 : 32;
        const int first_inval = inval_mask ? __ffs(inval_mask) - 1 : 32;
        if (first_inval < first_incl) {
          if (++spins > (1u << 24)) { if (lane == 0) atomicOr(p.err, DEV_ERR_WATCHDOG); break; }
          __nanosleep(32);
          continue;
        }
        int64_t val = lane <= first_incl ? (int64_t)(v & ST_VAL_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (first_incl < 32) break;
        j -= 32;
      }
      if (lane == 0) st_relaxed_u64(my_status + t, ST_FLAG_INCL | (uint64_t)(excl + total));
    }
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  const int64_t excl = s_excl;
  const int64_t e_base = s_hdr.e_in + excl;
  const int64_t s_base = s_hdr.s_in + excl;
  if (is_last && tid == 0) {
    p.e_len_out[b] = e_base + total;
    p.src_len_out[b] = s_base + total;
  }
  if (e_base + total > p.e_stride || s_base + total > p.src_stride) {
    if (tid == 0) atomicOr(p.err, DEV_ERR_CAPACITY);
    return;
  }
  if (total == 0) return;

  // ---- pass 2: decisions, one warp per node; q[s] = position AMONG THE PASSING EDGES chosen for slot s ----------
  const uint32_t pos0 = (uint32_t)(fb + node0);
  const uint32_t batch = p.batch_base + (uint32_t)b;
  uint8_t* s_owner = reinterpret_cast<uint8_t*>(s_mask + FT_THREADS * FT_MASKW);  // [tile_edges] output edge -> node of the tile
  constexpr uint32_t RESOLVED = 0x80000000u;
  // light nodes: the node's thread makes the decisions serially on its own slots (no atomics) and selects the chosen
  // edges from its two mask words
  if (tid < nn && s_deg[tid] <= FT_LIGHT) {
    const int n = tid;
    const uint32_t np = s_pass[n], o = s_off[n];
    const uint32_t c_n = KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE ? (np > 0 ? k : 0u) : min(np, k);
    if (c_n > 0) {
      uint32_t* q = s_slot + o;
      const uint32_t lo = s_mask[n * FT_MASKW], hi = s_mask[n * FT_MASKW + 1], nlo = __popc(lo);
      auto nth_passing = [&](uint32_t r) { return r < nlo ? __fns(lo, 0, (int)r + 1) : 32u + __fns(hi, 0, (int)(r - nlo) + 1); };
      for (uint32_t s2 = 0; s2 < c_n; ++s2) {
        q[s2] = 0u;
        s_owner[o + s2] = (uint8_t)n;
      }
      if (KIND == TCHGEO_SAMPLER_UNIFORM && np > k) {
        const uint32_t nb = (np - k + 3u) >> 2;
        for (uint32_t c = 0; c < nb; ++c) {          // sampling.rs:17-23 in step order: the last hit of a slot wins
          const Philox4 r = philox4x32_10(pos0 + n, c, batch, TAG_RESERVOIR | (p.rel << 8), p.key0, p.key1);
#pragma unroll
          for (uint32_t u = 0; u < 4; ++u) {
            const uint32_t step = k + 4u * c + u;
            const uint32_t jj = __umulhi(pick4(r, u), step);
            if (step < np && jj < k) q[jj] = step;
          }
        }
      } else if (KIND == TCHGEO_SAMPLER_WEIGHTED && np > k) {
        // w_sum in the reference's own order (w_sum = w_sum + w over the passing edges, sampling.rs:37-48)
        double w_sum = 0.0;
        const int64_t start = s_start[n];
        for (uint32_t rank = 0; rank < np; ++rank) {
          const double wv = __ldg(p.weights + start + nth_passing(rank));
          w_sum = w_sum + wv;
          if (rank < k) continue;
          if (!(w_sum > 0.0)) {
            atomicOr(p.err, DEV_ERR_PANIC);
            continue;
          }
          const Philox4 r = philox4x32_10(pos0 + n, rank, batch, TAG_WEIGHTED | (p.rel << 8), p.key0, p.key1);
          const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
          const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
          if (__dmul_rn(u, w_sum) < wv) q[__umulhi(r.z, k)] = rank;
        }
      }
      for (uint32_t s2 = 0; s2 < c_n; ++s2) {
        uint32_t r;
        if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
          const Philox4 rr = philox4x32_10(pos0 + n, s2 >> 2, batch, TAG_REPLACE | (p.rel << 8), p.key0, p.key1);
          r = __umulhi(pick4(rr, s2 & 3u), np);
        } else {
          r = q[s2] ? q[s2] : s2;
        }
        q[s2] = RESOLVED | nth_passing(r);
      }
    }
  }
  // hubs: one warp per node
  for (uint32_t hb = warp; hb < s_nhub; hb += NWARP) {
    const int n = s_hub[hb];
    const uint32_t np = s_pass[n], deg = s_deg[n], o = s_off[n];
    const uint32_t c_n = KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE ? (np > 0 ? k : 0u) : min(np, k);
    if (c_n == 0) continue;
    const int64_t start = s_start[n], state = s_state[n];
    uint32_t* q = s_slot + o;
    for (uint32_t s = lane; s < c_n; s += 32) {
      q[s] = 0u;
      s_owner[o + s] = (uint8_t)n;
    }
    __syncwarp();
    if (KIND == TCHGEO_SAMPLER_UNIFORM && np > k) {
      const uint32_t nb = (np - k + 3u) >> 2;
      for (uint32_t c = lane; c < nb; c += 32)
        reservoir_block(philox4x32_10(pos0 + n, c, batch, TAG_RESERVOIR | (p.rel << 8), p.key0, p.key1), k + 4u * c, np, k, q);
    } else if (KIND == TCHGEO_SAMPLER_WEIGHTED && np > k) {
      double carry = 0.0;
      uint32_t rank0 = 0;
      for (uint32_t base = 0; base < deg; base += 32) {
        const uint32_t item = base + lane;
        const bool ok = item < deg && filter_pass<STATIC>(p, __ldg(p.timestamps + start + item), state);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        const uint32_t rank = rank0 + __popc(m & ((1u << lane) - 1u));
        const double w = ok ? __ldg(p.weights + start + item) : 0.0;
        double incl = w;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, incl, sft);
          if (lane >= sft) incl += up;
        }
        const double w_sum = carry + incl;
        if (ok && rank >= k) {
          if (!(w_sum > 0.0)) atomicOr(p.err, DEV_ERR_PANIC);
          else {
            const Philox4 r = philox4x32_10(pos0 + n, rank, batch, TAG_WEIGHTED | (p.rel << 8), p.key0, p.key1);
            const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
            const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
            if (__dmul_rn(u, w_sum) < w) atomicMax(q + __umulhi(r.z, k), rank);
          }
        }
        carry = __shfl_sync(0xffffffffu, w_sum, 31);
        rank0 += __popc(m);
      }
    }
    __syncwarp();
    for (uint32_t s = lane; s < c_n; s += 32) {
      if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
        const Philox4 r = philox4x32_10(pos0 + n, s >> 2, batch, TAG_REPLACE | (p.rel << 8), p.key0, p.key1);
        q[s] = __umulhi(pick4(r, s & 3u), np);
      } else {
        const uint32_t st = q[s];
        q[s] = st ? st : s;
      }
    }
    __syncwarp();
    // select: passing position -> position inside the column.  One sweep over the timestamps; a chunk's ballot mask
    // tells which passing positions it holds, and the r-th set bit (__fns) is the edge: O(deg/32) per 32 slots instead
    // of the O(deg * k) match of every passing edge against every slot
    uint32_t rank0 = 0;
    for (uint32_t base = 0; base < deg && rank0 < np; base += 32) {
      uint32_t m;
      if (base < 32u * FT_MASKW) {
        m = s_mask[n * FT_MASKW + (base >> 5)];   // written by this warp's lane 0 in pass 1 (a barrier lies in between)
      } else {
        const uint32_t item = base + lane;
        const bool ok = item < deg && filter_pass<STATIC>(p, __ldg(p.timestamps + start + item), state);
        m = __ballot_sync(0xffffffffu, ok);
      }
      const uint32_t cm = __popc(m);
      for (uint32_t s = lane; s < c_n; s += 32) {
        const uint32_t r = q[s];
        if (!(r & RESOLVED) && r >= rank0 && r < rank0 + cm) q[s] = RESOLVED | (base + __fns(m, 0, (int)(r - rank0) + 1));
      }
      rank0 += cm;
    }
  }
  __syncthreads();

  // ---- emission: one thread per output edge, coalesced stores (the layout of neighbor_sampling.rs:210-218) ----------
  int64_t* o_s = p.src_samples + (int64_t)b * p.src_stride + s_base;
  int64_t* o_st = p.src_states + (int64_t)b * p.src_stride + s_base;
  int64_t* o_r = p.rows + (int64_t)b * p.e_stride + e_base;
  int64_t* o_c = p.cols + (int64_t)b * p.e_stride + e_base;
  int64_t* o_e = p.eidx + (int64_t)b * p.e_stride + e_base;
#pragma unroll 1
  for (uint32_t e = tid; e < total; e += FT_THREADS) {
    const uint32_t n = s_owner[e];
    const int64_t ptr = s_start[n] + (int64_t)(s_slot[e] & ~RESOLVED);
    const int64_t v = p.indices32 ? (int64_t)ld_gather64_i32(p.indices32 + ptr) : ld_gather64_i64(p.indices + ptr);
    const int64_t st_out = p.filter_mode == 3 ? __ldg(p.timestamps + ptr) : s_state[n];  // mutate(), :69-76
    st_cs_i64(o_s + e, v);
    st_cs_i64(o_st + e, st_out);
    st_cs_i64(o_r + e, s_base + e);
    st_cs_i64(o_c + e, fb + node0 + n);
    st_cs_i64(o_e + e, ptr);
  }
}

// ---- derived-array kernels of the graph handle ------------------------------------------------
// (weight, serial prefix sum) records of one relation: one thread per column accumulates in CSC order, exactly the
// reference's `w_sum = w_sum + w` sequence (sampling.rs:37-48; = csc_edge_cumsum, transform.rs:36-60).  One-time work.
__global__ void __launch_bounds__(256) wrec_kernel(const int64_t* __restrict__ ptrs, int64_t n_cols,
                                                   const double* __restrict__ w, int64_t nnz, double2* __restrict__ rec) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  int64_t s = ptrs[c], e = ptrs[c + 1];
  if (s < 0) s = 0;
  if (e > nnz) e = nnz;
  double acc = 0.0;
  for (int64_t q = s; q < e; ++q) {
    const double x = w[q];
    acc = acc + x;
    rec[q] = make_double2(x, acc);
  }
}

// Tuning knob (the default is what bench.py measures): TCHGEO_HOP_MIN_BLOCKS = 8 | 10 | 12 CTAs per SM the register
// budget is set for (64 / 48 / 40 registers).
inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
inline int hop_min_blocks() {
  static int v = -1;
  if (v < 0) {
    const int x = env_int("TCHGEO_HOP_MIN_BLOCKS", HOP_DEFAULT_MIN_BLOCKS);
    v = (x == 8 || x == 10 || x == 12) ? x : HOP_DEFAULT_MIN_BLOCKS;
  }
  return v;
}

}  // namespace
}  // namespace tchgeo

// ---- graph handle (struct tchgeo_graph: graph.cuh) ---------------------------------------------

namespace tchgeo {

// Builds the derived arrays named by `what` that are still missing.  Synchronises `stream` when it builds something.
tchgeo_status graph_ensure(tchgeo_graph* g, int32_t what, cudaStream_t stream) {
  TCHGEO_REQUIRE(g != nullptr, "graph is NULL");
  if ((what & TCHGEO_PREPARE_INDEX_REPLICA) && env_int("TCHGEO_INDEX_REPLICA", 1) != 0) {
    uint32_t* d_err = nullptr;
    for (int r = 0; r < g->R; ++r) {
      if (g->replica_state[(size_t)r] != 0) continue;
      g->replica_state[(size_t)r] = 2;
      const int64_t n = g->nnz[(size_t)r];
      if (n <= 0 || !g->indices[(size_t)r]) continue;
      if (!d_err) TCHGEO_CUDA_CHECK(cudaMalloc(&d_err, 8));
      int32_t* dst = nullptr;
      TCHGEO_CUDA_CHECK(cudaMalloc(&dst, (size_t)n * 4));
      uint32_t h[2] = {0u, 0u};
      tchgeo_status st = TCHGEO_OK;
      {
        cudaError_t e = cudaMemsetAsync(d_err, 0, 8, stream);
        if (e == cudaSuccess) {
          compress_kernel<true><<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(
              g->indices[(size_t)r], n, dst, d_err);
          e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_err, 8, cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
          cudaFree(dst);
          cudaFree(d_err);
          TCHGEO_CUDA_CHECK(e);
        }
        st = status_from_dev_err(h[0]);
      }
      if (st == TCHGEO_OK) {
        g->indices32[(size_t)r] = dst;
        g->replica_state[(size_t)r] = 1;
        g->max_index[(size_t)r] = (int64_t)h[1];
        g->derived_bytes += (size_t)n * 4;
      } else {
        cudaFree(dst);  // ids beyond int32 (or negative): sample from the i64 array; the kernels report bad ids
        if (st != TCHGEO_ERR_INDEX) {
          cudaFree(d_err);
          return st;
        }
      }
    }
    if (d_err) cudaFree(d_err);
  }
  if ((what & TCHGEO_PREPARE_WEIGHT_RECORDS) && env_int("TCHGEO_WEIGHT_CUMSUM", 1) != 0) {
    bool built = false;
    for (int r = 0; r < g->R; ++r) {
      if (g->wrec[(size_t)r] || !g->weights[(size_t)r] || g->nnz[(size_t)r] <= 0 || !g->ptrs[(size_t)r]) continue;
      double2* rec = nullptr;
      TCHGEO_CUDA_CHECK(cudaMalloc(&rec, (size_t)g->nnz[(size_t)r] * 16));
      const int64_t nc = g->num_major[(size_t)r];
      if (nc > 0) {
        wrec_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, stream>>>(g->ptrs[(size_t)r], nc, g->weights[(size_t)r],
                                                                     g->nnz[(size_t)r], rec);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
          cudaFree(rec);
          TCHGEO_CUDA_CHECK(e);
        }
      }
      g->wrec[(size_t)r] = rec;
      g->derived_bytes += (size_t)g->nnz[(size_t)r] * 16;
      built = true;
    }
    if (built) TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  }
  return TCHGEO_OK;
}

}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_graph_create(int32_t num_rels, const int64_t* const* ptrs, const int64_t* num_major,
                                             const int64_t* const* indices, const int64_t* nnz, tchgeo_graph_t** out) {
  TCHGEO_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  TCHGEO_REQUIRE(num_rels > 0 && ptrs && num_major && indices && nnz, "bad graph argument");
  for (int r = 0; r < num_rels; ++r)
    TCHGEO_REQUIRE(num_major[r] >= 0 && nnz[r] >= 0 && (nnz[r] == 0 || indices[r]), "relation %d: bad sizes", r);
  tchgeo_graph* g = new tchgeo_graph();
  const size_t R = (size_t)num_rels;
  g->R = num_rels;
  TCHGEO_CUDA_CHECK(cudaGetDevice(&g->device));
  g->ptrs.assign(ptrs, ptrs + R);
  g->indices.assign(indices, indices + R);
  g->num_major.assign(num_major, num_major + R);
  g->nnz.assign(nnz, nnz + R);
  g->weights.assign(R, nullptr);
  g->timestamps.assign(R, nullptr);
  g->indices32.assign(R, nullptr);
  g->replica_state.assign(R, 0);
  g->max_index.assign(R, -1);
  g->wrec.assign(R, nullptr);
  *out = g;
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_graph_set_weights(tchgeo_graph_t* g, const double* const* weights) {
  TCHGEO_REQUIRE(g != nullptr, "graph is NULL");
  for (int r = 0; r < g->R; ++r) {
    if (g->wrec[(size_t)r]) {
      cudaFree(g->wrec[(size_t)r]);
      g->derived_bytes -= (size_t)g->nnz[(size_t)r] * 16;
      g->wrec[(size_t)r] = nullptr;
    }
    g->weights[(size_t)r] = weights ? weights[r] : nullptr;
  }
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_graph_set_timestamps(tchgeo_graph_t* g, const int64_t* const* timestamps) {
  TCHGEO_REQUIRE(g != nullptr, "graph is NULL");
  for (int r = 0; r < g->R; ++r) g->timestamps[(size_t)r] = timestamps ? timestamps[r] : nullptr;
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_graph_prepare(tchgeo_graph_t* g, int32_t what, tchgeo_stream stream) {
  return graph_ensure(g, what, (cudaStream_t)stream);
}

extern "C" size_t tchgeo_graph_derived_bytes(const tchgeo_graph_t* g) { return g ? g->derived_bytes : 0; }

extern "C" void tchgeo_graph_destroy(tchgeo_graph_t* g) {
  if (!g) return;
  for (auto p : g->indices32) if (p) cudaFree(p);
  for (auto p : g->wrec) if (p) cudaFree(p);
  delete g;
}

// ---- host-side plan: which launches, which version rows of the length table ------------------
namespace tchgeo {
// csrc/relabel.cu
size_t relabel_workspace_bytes(int64_t num_trees, int64_t n_max, int64_t id_bound, bool prefer_waves = false);
int relabel_launches(int64_t num_trees, int64_t n_max, int64_t id_bound);
tchgeo_status relabel_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees,
                              int64_t num_seeds, int64_t n_max, int64_t id_bound, int64_t* nodes, int64_t* local,
                              int64_t* nodes_len, void* workspace, size_t workspace_bytes, uint32_t* err,
                              cudaStream_t stream, bool prefer_waves = false);
namespace {

struct Launch {
  int rel, hop;
  int fr_begin_row, fr_end_row;
  int src_in_row, src_out_row;
  int e_in_row, e_out_row;
  int dst_len_row;
  int64_t fanout;
  int tile_nodes, tile_edges, tiles_per_batch;
  size_t status_off;  // in uint64 words
};

// effective per-relation device arrays of a call: the tables of the args, or those of its graph handle
struct Resolved {
  std::vector<const int64_t*> ptrs, indices, timestamps;
  std::vector<int64_t> num_cols, nnz;
  std::vector<const double*> weights;
  std::vector<const int32_t*> indices32;
  std::vector<const double2*> wrec;
  std::vector<int64_t> max_index;   // largest id of indices[r], -1 = unknown (no graph handle / no replica)
};

struct Plan {
  int T, R, H;
  int64_t B;
  std::vector<Launch> launches;
  std::vector<int> n_row0;              // [T] initial length rows
  std::vector<int> n_row_final;         // [T]
  std::vector<int> e_row_final;         // [R]
  std::vector<int> lo_rows;             // [R*H*3] rows for layer offsets, -1 = relation inactive
  std::vector<int64_t> samples_cap, edges_cap;  // worst case per batch
  int num_rows;                         // V
  size_t status_words;
  // dedup + relabel stage (K7): one length row per node type; id_bound = what is known about the type's ids
  // (0: any i64; 2^32-1: they fit 32 bits; less: the type's node count, which lets the stage use direct-address tables)
  bool relabel;
  std::vector<int> nodes_row;           // [T] or -1
  std::vector<int64_t> id_bound;        // [T]
  int relabel_kernels;
  // workspace layout (bytes)
  size_t off_ctrl, off_state, off_status, off_relabel, relabel_bytes, total_bytes;
  Resolved rs;
};

constexpr size_t CTRL_WORDS = 64;  // uint32: [0] = err, [1..] = one ticket per launch (grown if needed)

tchgeo_status resolve(const tchgeo_sampling_args* a, Resolved& rs, bool want_derived, cudaStream_t stream) {
  const size_t R = (size_t)a->num_rels;
  rs.ptrs.assign(R, nullptr); rs.indices.assign(R, nullptr); rs.timestamps.assign(R, nullptr);
  rs.num_cols.assign(R, 0); rs.nnz.assign(R, INT64_MAX);
  rs.weights.assign(R, nullptr); rs.indices32.assign(R, nullptr); rs.wrec.assign(R, nullptr);
  rs.max_index.assign(R, -1);
  const bool weighted = a->sampler_kind == TCHGEO_SAMPLER_WEIGHTED;
  if (a->graph) {
    tchgeo_graph* g = const_cast<tchgeo_graph*>(a->graph);
    TCHGEO_REQUIRE(g->R == a->num_rels, "graph handle has %d relations, the call %d", g->R, a->num_rels);
    if (want_derived) {
      int what = TCHGEO_PREPARE_INDEX_REPLICA;
      if (weighted && !a->filter_mode) what |= TCHGEO_PREPARE_WEIGHT_RECORDS;  // with a filter w_sum runs over the passing edges
      const tchgeo_status st = graph_ensure(g, what, stream);
      if (st != TCHGEO_OK) return st;
    }
    for (size_t r = 0; r < R; ++r) {
      rs.ptrs[r] = g->ptrs[r]; rs.indices[r] = g->indices[r]; rs.num_cols[r] = g->num_major[r]; rs.nnz[r] = g->nnz[r];
      rs.weights[r] = weighted ? g->weights[r] : nullptr;
      rs.timestamps[r] = (a->timestamps && a->timestamps[r]) ? a->timestamps[r] : g->timestamps[r];
      rs.indices32[r] = env_int("TCHGEO_INDEX_REPLICA", 1) != 0 ? g->indices32[r] : nullptr;
      rs.max_index[r] = g->max_index[r];
      rs.wrec[r] = (weighted && !a->filter_mode) ? g->wrec[r] : nullptr;
    }
    return TCHGEO_OK;
  }
  if (!(a->col_ptrs && a->row_indices && a->num_cols)) return TCHGEO_OK;  // geometry queries; check_args refuses to run
  for (size_t r = 0; r < R; ++r) {
    rs.ptrs[r] = a->col_ptrs[r]; rs.indices[r] = a->row_indices[r]; rs.num_cols[r] = a->num_cols[r];
    if (a->nnz) rs.nnz[r] = a->nnz[r];
    if (weighted && a->weights) rs.weights[r] = a->weights[r];
    if (a->filter_mode && a->timestamps) rs.timestamps[r] = a->timestamps[r];
  }
  return TCHGEO_OK;
}

// geometry only (capacities, rows, launches): needs no device arrays
tchgeo_status build_plan(const tchgeo_sampling_args* a, Plan& pl) {
  TCHGEO_REQUIRE(a != nullptr, "args is NULL");
  TCHGEO_REQUIRE(a->num_node_types > 0 && a->num_rels > 0 && a->num_hops >= 0, "bad T/R/H");
  TCHGEO_REQUIRE(a->num_batches > 0 && a->num_batches < (1ll << 24), "num_batches out of range");
  TCHGEO_REQUIRE(a->rel_src && a->rel_dst && (a->fanouts || a->num_hops == 0) && a->seeds_per_batch, "NULL host array");
  TCHGEO_REQUIRE(a->sampler_kind >= 0 && a->sampler_kind <= 2, "unknown sampler kind");
  pl.T = a->num_node_types; pl.R = a->num_rels; pl.H = a->num_hops; pl.B = a->num_batches;
  const int T = pl.T, R = pl.R, H = pl.H;
  for (int r = 0; r < R; ++r) {
    TCHGEO_REQUIRE(a->rel_src[r] >= 0 && a->rel_src[r] < T && a->rel_dst[r] >= 0 && a->rel_dst[r] < T,
                   "relation %d: node type index out of range", r);
    for (int h = 0; h < H; ++h) {
      const int64_t k = a->fanouts[(size_t)r * H + h];
      TCHGEO_REQUIRE(k >= 0 && k <= MAX_FANOUT, "relation %d hop %d: fanout %lld unsupported (max %d)", r, h,
                     (long long)k, MAX_FANOUT);
    }
  }
  int rows = 1;  // row 0 = zeros
  pl.n_row0.assign(T, 0);
  std::vector<int> cur_n(T), sl_begin(T), sl_end(T), cur_e(R, 0);
  std::vector<int64_t> wc_len(T), wc_front(T);
  pl.samples_cap.assign(T, 0);
  pl.edges_cap.assign(R, 0);
  for (int t = 0; t < T; ++t) {
    TCHGEO_REQUIRE(a->seeds_per_batch[t] >= 0, "negative seed count");
    pl.n_row0[t] = rows++;
    cur_n[t] = pl.n_row0[t];
    sl_begin[t] = 0;
    sl_end[t] = cur_n[t];
    wc_len[t] = wc_front[t] = a->seeds_per_batch[t];
  }
  pl.lo_rows.assign((size_t)R * std::max(H, 1) * 3, -1);
  pl.launches.clear();
  size_t status_words = 0;
  const int64_t LIM = (int64_t)1 << 40;
  for (int h = 0; h < H; ++h) {
    std::vector<int64_t> add(T, 0);
    for (int r = 0; r < R; ++r) {
      if (a->rel_active && !a->rel_active[r]) continue;
      const int st = a->rel_src[r], dt = a->rel_dst[r];
      const int64_t k = a->fanouts[(size_t)r * H + h];
      int* lo = &pl.lo_rows[((size_t)r * H + h) * 3];
      lo[0] = cur_n[st]; lo[1] = cur_e[r]; lo[2] = cur_n[dt];
      const int64_t fcap = wc_front[dt];
      if (fcap == 0 || k == 0) {
        // nothing can be appended: lengths keep their current version rows.
        // (k == 0 with a non-empty neighbourhood panics in the reference unless sampling with
        //  replacement; that case is detected by a launch below only when k > 0, so flag it here.)
        if (k == 0 && fcap > 0 && a->sampler_kind != TCHGEO_SAMPLER_UNIFORM_REPLACE) {
          // keep the launch so the kernel can raise DEV_ERR_PANIC for deg > 0
        } else {
          continue;
        }
      }
      Launch L;
      L.rel = r; L.hop = h; L.fanout = k;
      L.fr_begin_row = sl_begin[dt]; L.fr_end_row = sl_end[dt];
      L.src_in_row = cur_n[st]; L.src_out_row = rows++;
      L.e_in_row = cur_e[r]; L.e_out_row = rows++;
      L.dst_len_row = cur_n[dt];
      const int64_t kk = std::max<int64_t>(k, 1);
      const int tile_threads = a->filter_mode ? FT_THREADS : HOP_THREADS;
      int tn = (int)std::min<int64_t>(tile_threads, std::max<int64_t>(1, MAX_TILE_EDGES / kk));
      L.tile_nodes = tn;
      L.tile_edges = (int)(tn * kk);
      const int64_t tpb = (fcap + tn - 1) / tn;
      TCHGEO_REQUIRE(tpb * pl.B < ((int64_t)1 << 31), "grid too large: %lld tiles x %lld batches", (long long)tpb,
                     (long long)pl.B);
      L.tiles_per_batch = (int)tpb;
      L.status_off = status_words;
      status_words += (size_t)tpb * (size_t)pl.B;
      pl.launches.push_back(L);
      cur_n[st] = L.src_out_row;
      cur_e[r] = L.e_out_row;
      const int64_t n_new = fcap * k;
      TCHGEO_REQUIRE(n_new < LIM && wc_len[st] + n_new < LIM, "worst-case output too large");
      pl.edges_cap[r] += n_new;
      add[st] += n_new;
      wc_len[st] += n_new;
    }
    for (int t = 0; t < T; ++t) {  // neighbor_sampling.rs:345-348
      sl_begin[t] = sl_end[t];
      sl_end[t] = cur_n[t];
      wc_front[t] = add[t];
    }
  }
  for (int t = 0; t < T; ++t) pl.samples_cap[t] = wc_len[t];
  pl.n_row_final = cur_n;
  pl.e_row_final = cur_e;
  pl.relabel = a->nodes != nullptr || a->local != nullptr;
  pl.nodes_row.assign(T, -1);
  if (pl.relabel)
    for (int t = 0; t < T; ++t) pl.nodes_row[t] = rows++;
  pl.num_rows = rows;
  pl.status_words = status_words;
  const size_t ctrl_words = std::max(CTRL_WORDS, pl.launches.size() + 2);
  pl.off_ctrl = 0;
  pl.off_state = ((ctrl_words * 4 + 255) / 256) * 256;
  pl.off_status = pl.off_state + (((size_t)rows * pl.B * 8 + 255) / 256) * 256;
  pl.off_relabel = pl.off_status + ((status_words * 8 + 255) / 256) * 256;
  pl.relabel_bytes = 0;
  pl.relabel_kernels = 0;
  pl.id_bound.assign(T, 0);
  pl.total_bytes = pl.off_relabel + 256;
  return TCHGEO_OK;
}

// everything: geometry + the device arrays the launches read + the relabel stage's share of the workspace
tchgeo_status make_plan(const tchgeo_sampling_args* a, Plan& pl, bool want_derived) {
  tchgeo_status st = build_plan(a, pl);
  if (st != TCHGEO_OK) return st;
  st = resolve(a, pl.rs, want_derived, (cudaStream_t)a->stream);
  if (st != TCHGEO_OK) return st;
  if (pl.relabel) {
    TCHGEO_REQUIRE(a->nodes && a->local, "relabel needs both nodes and local");
    for (int t = 0; t < pl.T; ++t) {
      // 32-bit keys when every id of the type is known to be below 2^31: all relations that append to it read an
      // int32 replica (seeds beyond the bound are then reported as TCHGEO_ERR_INDEX by the stage).  The bound itself:
      // the type's node count as far as the graph tells -- the columns of the relations it is the destination of (its
      // seeds are checked against those by the hop kernels) and the largest source id of the relations that append to it.
      bool k32 = true, fed = false, known = true, is_dst = false;
      int64_t bound = 0;
      for (int r = 0; r < pl.R; ++r) {
        if (a->rel_active && !a->rel_active[r]) continue;
        if (a->rel_src[r] == t) {
          fed = true;
          k32 = k32 && pl.rs.indices32[(size_t)r] != nullptr;
          if (pl.rs.max_index[(size_t)r] < 0) known = false;
          else bound = std::max(bound, pl.rs.max_index[(size_t)r] + 1);
        }
        if (a->rel_dst[r] == t) {
          is_dst = true;
          bound = std::max(bound, pl.rs.num_cols[(size_t)r]);
        }
      }
      if (a->seeds_per_batch[t] > 0 && !is_dst) known = false;   // nothing checks those seeds
      pl.id_bound[t] = !(k32 && fed) ? 0 : (known && bound > 0 && bound < 0xFFFFFFFFll ? bound : 0xFFFFFFFFll);
      if (pl.samples_cap[t] == 0) continue;
      const size_t need = relabel_workspace_bytes(pl.B, pl.samples_cap[t], pl.id_bound[t]);
      TCHGEO_REQUIRE(need > 0, "relabel: tree of node type %d too large", t);
      pl.relabel_bytes = std::max(pl.relabel_bytes, need);
      pl.relabel_kernels += relabel_launches(pl.B, pl.samples_cap[t], pl.id_bound[t]);
    }
    pl.total_bytes = pl.off_relabel + pl.relabel_bytes + 256;
  }
  return TCHGEO_OK;
}

template <int KIND, int MINB, bool I32>
cudaError_t launch_hop_v(const HopParams& hp, int64_t tiles, size_t smem, cudaStream_t stream) {
  static bool configured[64] = {};  // per device; benign race: the attribute is idempotent
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(hop_kernel<KIND, MINB, I32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int64_t grid = (tiles + HOP_TPC - 1) / HOP_TPC;
  hop_kernel<KIND, MINB, I32><<<(unsigned)grid, HOP_THREADS, smem, stream>>>(hp);
  return cudaGetLastError();
}

template <int KIND, bool I32>
cudaError_t launch_hop_i(const HopParams& hp, int64_t tiles, size_t smem, cudaStream_t stream) {
  switch (hop_min_blocks()) {
    case 8: return launch_hop_v<KIND, 8, I32>(hp, tiles, smem, stream);
    case 12: return launch_hop_v<KIND, 12, I32>(hp, tiles, smem, stream);
    default: return launch_hop_v<KIND, 10, I32>(hp, tiles, smem, stream);
  }
}

template <int KIND>
cudaError_t launch_hop(const HopParams& hp, int64_t tiles, size_t smem, cudaStream_t stream) {
  return hp.indices32 ? launch_hop_i<KIND, true>(hp, tiles, smem, stream)
                      : launch_hop_i<KIND, false>(hp, tiles, smem, stream);
}

template <int KIND, bool STATIC>
cudaError_t launch_filtered_m(const HopParams& hp, int64_t grid, size_t smem, cudaStream_t stream) {
  static bool configured[64] = {};  // per device: fanouts above ~9 k need more than the default 48 KB
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(hop_filtered_kernel<KIND, STATIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  hop_filtered_kernel<KIND, STATIC><<<(unsigned)grid, FT_THREADS, smem, stream>>>(hp);
  return cudaGetLastError();
}
template <int KIND>
cudaError_t launch_filtered(const HopParams& hp, int64_t grid, size_t smem, cudaStream_t stream) {
  return hp.filter_mode == 1 ? launch_filtered_m<KIND, true>(hp, grid, smem, stream)
                             : launch_filtered_m<KIND, false>(hp, grid, smem, stream);
}

struct EventList {  // destroyed on every exit path
  std::vector<cudaEvent_t> ev;
  ~EventList() {
    for (auto e : ev) cudaEventDestroy(e);
  }
};

// decode a host copy of the length table into the result arrays (any may be NULL)
void decode_state(const Plan& pl, const int64_t* state, int64_t* samples_len, int64_t* edges_len, int64_t* layer_offsets,
                  int64_t* nodes_len) {
  const int T = pl.T, R = pl.R, H = pl.H;
  const int64_t B = pl.B;
  for (int64_t b = 0; b < B; ++b) {
    if (samples_len)
      for (int t = 0; t < T; ++t) samples_len[b * T + t] = state[(size_t)pl.n_row_final[t] * B + b];
    if (edges_len)
      for (int r = 0; r < R; ++r) edges_len[b * R + r] = state[(size_t)pl.e_row_final[r] * B + b];
    if (layer_offsets)
      for (int r = 0; r < R; ++r)
        for (int h = 0; h < H; ++h)
          for (int c = 0; c < 3; ++c) {
            const int row = pl.lo_rows[((size_t)r * H + h) * 3 + c];
            layer_offsets[(((size_t)b * R + r) * H + h) * 3 + c] = row < 0 ? -1 : state[(size_t)row * B + b];
          }
    if (nodes_len && pl.relabel)
      for (int t = 0; t < T; ++t) nodes_len[b * T + t] = pl.nodes_row[t] < 0 ? 0 : state[(size_t)pl.nodes_row[t] * B + b];
  }
}

tchgeo_status check_args(const tchgeo_sampling_args* a, const Plan& pl) {
  const int T = pl.T;
  TCHGEO_REQUIRE(a->inputs && a->samples && a->samples_stride && a->rows && a->cols && a->edge_index && a->edges_stride,
                 "NULL pointer table");
  TCHGEO_REQUIRE(a->workspace != nullptr, "workspace is NULL");
  if (a->workspace_bytes < pl.total_bytes) {
    set_last_error("workspace too small: need %zu bytes, got %zu", pl.total_bytes, a->workspace_bytes);
    return TCHGEO_ERR_CAPACITY;
  }
  for (int t = 0; t < T; ++t) {
    if (a->seeds_per_batch[t] > 0) TCHGEO_REQUIRE(a->inputs[t] != nullptr, "inputs[%d] is NULL", t);
    if (a->samples_stride[t] < a->seeds_per_batch[t]) {
      set_last_error("samples_stride[%d] smaller than the seed count", t);
      return TCHGEO_ERR_CAPACITY;
    }
    if (a->samples_stride[t] > 0) TCHGEO_REQUIRE(a->samples[t] != nullptr, "samples[%d] is NULL", t);
    TCHGEO_REQUIRE(a->samples_stride[t] < ((int64_t)1 << 32), "samples_stride[%d] must be < 2^32", t);
    if (pl.relabel && a->samples_stride[t] > 0)
      TCHGEO_REQUIRE(a->nodes[t] && a->local[t], "relabel: nodes[%d] / local[%d] is NULL", t, t);
  }
  for (const Launch& L : pl.launches) {
    const int r = L.rel;
    TCHGEO_REQUIRE(pl.rs.ptrs[(size_t)r] && pl.rs.num_cols[(size_t)r] >= 0, "relation %d: col_ptrs is NULL", r);
    // row_indices[r] may be NULL for a relation without edges: it is only dereferenced when deg > 0
    if (a->sampler_kind == TCHGEO_SAMPLER_WEIGHTED)
      TCHGEO_REQUIRE(pl.rs.weights[(size_t)r], "relation %d: weighted sampler without weights", r);
    if (a->filter_mode) TCHGEO_REQUIRE(pl.rs.timestamps[(size_t)r], "relation %d: temporal filter without timestamps", r);
    if (a->edges_stride[r] > 0)
      TCHGEO_REQUIRE(a->rows[r] && a->cols[r] && a->edge_index[r], "relation %d: NULL edge output", r);
  }
  if (a->filter_mode) {
    TCHGEO_REQUIRE(a->filter_mode >= 1 && a->filter_mode <= 3, "unknown filter mode");
    TCHGEO_REQUIRE(a->states && a->inputs_state, "temporal filter needs states and inputs_state");
    for (int t = 0; t < T; ++t) {
      if (a->samples_stride[t] > 0) TCHGEO_REQUIRE(a->states[t] != nullptr, "states[%d] is NULL", t);
      if (a->seeds_per_batch[t] > 0) TCHGEO_REQUIRE(a->inputs_state[t] != nullptr, "inputs_state[%d] is NULL", t);
    }
  }
  return TCHGEO_OK;
}

// Enqueues one sampling step on `stream`: no host synchronisation.  ev (optional): n_launch + 1 (+ 1 with relabel)
// events recorded around the launches.
tchgeo_status enqueue_step(const tchgeo_sampling_args* a, const Plan& pl, uint64_t seed, uint32_t batch_base,
                           cudaStream_t stream, std::vector<cudaEvent_t>* ev) {
  const int T = pl.T;
  const int64_t B = pl.B;
  char* ws = (char*)a->workspace;
  uint32_t* ctrl = (uint32_t*)(ws + pl.off_ctrl);
  int64_t* state = (int64_t*)(ws + pl.off_state);
  uint64_t* status = (uint64_t*)(ws + pl.off_status);

  // control words, length table and look-back status all start at zero
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws, 0, pl.off_relabel, stream));
  for (int t = 0; t < T; ++t) {
    const int64_t S = a->seeds_per_batch[t];
    if (S == 0) continue;
    // samples[t][b, 0:S] = inputs[t][b, :]   (neighbor_sampling.rs:184 / :264-271)
    TCHGEO_CUDA_CHECK(cudaMemcpy2DAsync(a->samples[t], (size_t)a->samples_stride[t] * 8, a->inputs[t], (size_t)S * 8,
                                        (size_t)S * 8, (size_t)B, cudaMemcpyDeviceToDevice, stream));
    fill_i64_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(state + (size_t)pl.n_row0[t] * B, S, B);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    if (a->filter_mode)  // states.extend_from_slice(inputs_state), neighbor_sampling.rs:185 / :273-277
      TCHGEO_CUDA_CHECK(cudaMemcpy2DAsync(a->states[t], (size_t)a->samples_stride[t] * 8, a->inputs_state[t],
                                          (size_t)S * 8, (size_t)S * 8, (size_t)B, cudaMemcpyDeviceToDevice, stream));
  }
  int li = 0;
  if (ev) TCHGEO_CUDA_CHECK(cudaEventRecord((*ev)[0], stream));
  for (const Launch& L : pl.launches) {
    const int r = L.rel, stt = a->rel_src[r], dtt = a->rel_dst[r];
    HopParams hp;
    hp.ptrs = pl.rs.ptrs[(size_t)r];
    hp.indices = pl.rs.indices[(size_t)r];
    hp.indices32 = pl.rs.indices32[(size_t)r];
    hp.weights = pl.rs.weights[(size_t)r];
    hp.wrec = pl.rs.wrec[(size_t)r];
    hp.num_cols = pl.rs.num_cols[(size_t)r];
    hp.nnz = pl.rs.nnz[(size_t)r];
    hp.dst_samples = a->samples[dtt];
    hp.dst_stride = a->samples_stride[dtt];
    hp.src_samples = a->samples[stt];
    hp.src_stride = a->samples_stride[stt];
    hp.rows = a->rows[r];
    hp.cols = a->cols[r];
    hp.eidx = a->edge_index[r];
    hp.e_stride = a->edges_stride[r];
    hp.fr_begin = state + (size_t)L.fr_begin_row * B;
    hp.fr_end = state + (size_t)L.fr_end_row * B;
    hp.src_len_in = state + (size_t)L.src_in_row * B;
    hp.src_len_out = state + (size_t)L.src_out_row * B;
    hp.e_len_in = state + (size_t)L.e_in_row * B;
    hp.e_len_out = state + (size_t)L.e_out_row * B;
    hp.status = status + L.status_off;
    hp.ticket = ctrl + 1 + li;
    hp.err = ctrl;
    hp.num_batches = (int32_t)B;
    hp.tiles_per_batch = L.tiles_per_batch;
    hp.tile_nodes = L.tile_nodes;
    hp.tile_edges = L.tile_edges;
    hp.fanout = (int32_t)L.fanout;
    hp.key0 = (uint32_t)seed;
    hp.key1 = (uint32_t)(seed >> 32);
    for (uint32_t q = 0; q < 10; ++q) {
      hp.rk[2 * q] = hp.key0 + q * 0x9E3779B9u;
      hp.rk[2 * q + 1] = hp.key1 + q * 0xBB67AE85u;
    }
    hp.rel = (uint32_t)r;
    hp.batch_base = batch_base;
    const int64_t grid = (int64_t)L.tiles_per_batch * B;
    const size_t smem = (size_t)L.tile_edges * 5 + 16;
    hp.total_tiles = (uint32_t)((int64_t)L.tiles_per_batch * B);
    hp.filter_mode = a->filter_mode;
    hp.filter_forward = a->filter_forward;
    hp.win_lo = a->filter_window_lo;
    hp.win_hi = a->filter_window_hi;
    hp.timestamps = a->filter_mode ? pl.rs.timestamps[(size_t)r] : nullptr;
    hp.dst_states = a->filter_mode ? a->states[dtt] : nullptr;
    hp.src_states = a->filter_mode ? a->states[stt] : nullptr;
    cudaError_t e;
    if (a->filter_mode) {
      const size_t fsmem = (size_t)L.tile_edges * 5 + (size_t)FT_THREADS * FT_MASKW * 4 + 16;  // positions, masks, owners
      switch (a->sampler_kind) {
        case TCHGEO_SAMPLER_UNIFORM: e = launch_filtered<TCHGEO_SAMPLER_UNIFORM>(hp, grid, fsmem, stream); break;
        case TCHGEO_SAMPLER_UNIFORM_REPLACE: e = launch_filtered<TCHGEO_SAMPLER_UNIFORM_REPLACE>(hp, grid, fsmem, stream); break;
        default: e = launch_filtered<TCHGEO_SAMPLER_WEIGHTED>(hp, grid, fsmem, stream); break;
      }
    } else {
      switch (a->sampler_kind) {
        case TCHGEO_SAMPLER_UNIFORM: e = launch_hop<TCHGEO_SAMPLER_UNIFORM>(hp, grid, smem, stream); break;
        case TCHGEO_SAMPLER_UNIFORM_REPLACE: e = launch_hop<TCHGEO_SAMPLER_UNIFORM_REPLACE>(hp, grid, smem, stream); break;
        default: e = launch_hop<TCHGEO_SAMPLER_WEIGHTED>(hp, grid, smem, stream); break;
      }
    }
    TCHGEO_CUDA_CHECK(e);
    ++li;
    if (ev) TCHGEO_CUDA_CHECK(cudaEventRecord((*ev)[(size_t)li], stream));
  }
  if (pl.relabel) {
    // K7: dedup + insertion-order relabel of every batch's tree, per node type, from the device-side lengths
    for (int t = 0; t < T; ++t) {
      const int64_t cap = std::min(pl.samples_cap[t], a->samples_stride[t]);
      if (cap <= 0) continue;
      const tchgeo_status st = relabel_enqueue(a->samples[t], a->samples_stride[t], state + (size_t)pl.n_row_final[t] * B, B,
                                               a->seeds_per_batch[t], cap, pl.id_bound[t], a->nodes[t], a->local[t],
                                               state + (size_t)pl.nodes_row[t] * B, ws + pl.off_relabel, pl.relabel_bytes + 256,
                                               ctrl, stream);
      if (st != TCHGEO_OK) return st;
    }
    if (ev) TCHGEO_CUDA_CHECK(cudaEventRecord((*ev)[(size_t)li + 1], stream));
  }
  return TCHGEO_OK;
}

tchgeo_status make_events(EventList& evs, size_t n) {
  evs.ev.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    cudaEvent_t e;
    TCHGEO_CUDA_CHECK(cudaEventCreate(&e));
    evs.ev.push_back(e);
  }
  return TCHGEO_OK;
}

int read_intervals(const EventList& evs, float* launch_ms) {
  const int n = (int)evs.ev.size() - 1;
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, evs.ev[(size_t)i], evs.ev[(size_t)i + 1]) != cudaSuccess) ms = -1.f;
    launch_ms[i] = ms;
  }
  return n;
}

tchgeo_status collect_with_plan(const tchgeo_sampling_args* a, const Plan& pl) {
  TCHGEO_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= pl.total_bytes, "workspace too small");
  cudaStream_t stream = (cudaStream_t)a->stream;
  char* ws = (char*)a->workspace;
  const size_t n_state = (size_t)pl.num_rows * pl.B;
  std::vector<int64_t> state(n_state);
  uint32_t err = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(state.data(), ws + pl.off_state, n_state * 8, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&err, ws + pl.off_ctrl, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  decode_state(pl, state.data(), a->samples_len, a->edges_len, a->layer_offsets, a->nodes_len);
  return status_from_dev_err(err);
}

tchgeo_status run_sampling(const tchgeo_sampling_args* a, float* launch_ms, int32_t launch_ms_cap, int32_t* num_launches) {
  Plan pl;
  tchgeo_status st = make_plan(a, pl, true);
  if (st != TCHGEO_OK) return st;
  st = check_args(a, pl);
  if (st != TCHGEO_OK) return st;
  const int n_iv = (int)pl.launches.size() + (pl.relabel ? 1 : 0);
  if (num_launches) *num_launches = n_iv;
  EventList evs;
  if (launch_ms) {
    TCHGEO_REQUIRE(launch_ms_cap >= n_iv, "launch_ms too small: %d intervals", n_iv);
    st = make_events(evs, (size_t)n_iv + 1);
    if (st != TCHGEO_OK) return st;
  }
  st = enqueue_step(a, pl, a->seed, a->batch_base, (cudaStream_t)a->stream, launch_ms ? &evs.ev : nullptr);
  if (st != TCHGEO_OK) return st;
  if (launch_ms || a->samples_len || a->edges_len || a->layer_offsets || a->nodes_len) st = collect_with_plan(a, pl);
  if (launch_ms) read_intervals(evs, launch_ms);
  return st;
}

}  // namespace
}  // namespace tchgeo

extern "C" tchgeo_status tchgeo_neighbor_sampling_capacity(const tchgeo_sampling_args* args, int64_t* samples_cap,
                                                           int64_t* edges_cap) {
  Plan pl;
  tchgeo_status st = build_plan(args, pl);
  if (st != TCHGEO_OK) return st;
  if (samples_cap) for (int t = 0; t < pl.T; ++t) samples_cap[t] = pl.samples_cap[t];
  if (edges_cap) for (int r = 0; r < pl.R; ++r) edges_cap[r] = pl.edges_cap[r];
  return TCHGEO_OK;
}

// (with a graph handle this builds the handle's derived arrays if they are missing: the relabel stage sizes its hash
//  slots by whether the int32 replica exists)
extern "C" size_t tchgeo_neighbor_sampling_workspace_bytes(const tchgeo_sampling_args* args) {
  Plan pl;
  if (make_plan(args, pl, true) != TCHGEO_OK) return 0;
  return pl.total_bytes;
}

extern "C" tchgeo_status tchgeo_neighbor_sampling_collect(const tchgeo_sampling_args* a) {
  Plan pl;
  tchgeo_status st = make_plan(a, pl, false);
  if (st != TCHGEO_OK) return st;
  return collect_with_plan(a, pl);
}

extern "C" tchgeo_status tchgeo_neighbor_sampling(const tchgeo_sampling_args* a) {
  return run_sampling(a, nullptr, 0, nullptr);
}

extern "C" tchgeo_status tchgeo_neighbor_sampling_timed(const tchgeo_sampling_args* a, float* launch_ms,
                                                        int32_t launch_ms_cap, int32_t* num_launches) {
  TCHGEO_REQUIRE(launch_ms != nullptr, "launch_ms is NULL");
  return run_sampling(a, launch_ms, launch_ms_cap, num_launches);
}

// ---- plan handle -------------------------------------------------------------------------------
struct tchgeo_plan {
  tchgeo_sampling_args a;  // deep copy: every HOST table below is owned by the handle
  std::vector<int32_t> rel_src, rel_dst;
  std::vector<const int64_t*> col_ptrs, row_indices, inputs, timestamps, inputs_state;
  std::vector<const double*> weights;
  std::vector<int64_t> num_cols, nnz, fanouts, seeds_per_batch, samples_stride, edges_stride;
  std::vector<uint8_t> rel_active;
  std::vector<int64_t*> samples, rows, cols, edge_index, states, nodes, local;
  Plan pl;
  int device = 0;
  int64_t* h_state = nullptr;  // pinned: length table + [n_state] = error word
  size_t n_state = 0;
  cudaEvent_t done = nullptr;
  bool pending = false;
  std::vector<int64_t> samples_len, edges_len, layer_offsets, nodes_len;
};

namespace {
template <typename V, typename P>
void copy_table(V& v, P*& field, size_t n) {
  if (field == nullptr || n == 0) {
    field = nullptr;
    return;
  }
  v.assign(field, field + n);
  field = v.data();
}
}  // namespace

extern "C" void tchgeo_plan_destroy(tchgeo_plan_t* p) {
  if (!p) return;
  if (p->done) cudaEventDestroy(p->done);
  if (p->h_state) cudaFreeHost(p->h_state);
  delete p;
}

extern "C" tchgeo_status tchgeo_plan_create(const tchgeo_sampling_args* args, tchgeo_plan_t** out) {
  TCHGEO_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  TCHGEO_REQUIRE(args != nullptr && args->num_node_types > 0 && args->num_rels > 0 && args->num_hops >= 0, "bad T/R/H");
  tchgeo_plan* p = new tchgeo_plan();
  p->a = *args;
  tchgeo_sampling_args& a = p->a;
  const size_t T = (size_t)a.num_node_types, R = (size_t)a.num_rels, H = (size_t)a.num_hops;
  copy_table(p->rel_src, a.rel_src, R);
  copy_table(p->rel_dst, a.rel_dst, R);
  copy_table(p->col_ptrs, a.col_ptrs, R);
  copy_table(p->num_cols, a.num_cols, R);
  copy_table(p->row_indices, a.row_indices, R);
  copy_table(p->weights, a.weights, R);
  copy_table(p->nnz, a.nnz, R);
  copy_table(p->fanouts, a.fanouts, R * H);
  copy_table(p->rel_active, a.rel_active, R);
  copy_table(p->inputs, a.inputs, T);
  copy_table(p->seeds_per_batch, a.seeds_per_batch, T);
  copy_table(p->samples, a.samples, T);
  copy_table(p->samples_stride, a.samples_stride, T);
  copy_table(p->rows, a.rows, R);
  copy_table(p->cols, a.cols, R);
  copy_table(p->edge_index, a.edge_index, R);
  copy_table(p->edges_stride, a.edges_stride, R);
  copy_table(p->timestamps, a.timestamps, R);
  copy_table(p->inputs_state, a.inputs_state, T);
  copy_table(p->states, a.states, T);
  copy_table(p->nodes, a.nodes, T);
  copy_table(p->local, a.local, T);
  a.samples_len = a.edges_len = a.layer_offsets = a.nodes_len = nullptr;
  tchgeo_status st = make_plan(&a, p->pl, true);
  if (st == TCHGEO_OK) st = check_args(&a, p->pl);
  if (st != TCHGEO_OK) {
    tchgeo_plan_destroy(p);
    return st;
  }
  const Plan& pl = p->pl;
  p->n_state = (size_t)pl.num_rows * pl.B;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMallocHost(&p->h_state, (p->n_state + 1) * 8);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    tchgeo_plan_destroy(p);
    TCHGEO_CUDA_CHECK(e);
  }
  p->samples_len.assign((size_t)pl.B * T, 0);
  p->edges_len.assign((size_t)pl.B * R, 0);
  p->layer_offsets.assign((size_t)pl.B * R * std::max<size_t>(H, 1) * 3, 0);
  if (pl.relabel) p->nodes_len.assign((size_t)pl.B * T, 0);
  *out = p;
  return TCHGEO_OK;
}

static tchgeo_status plan_enqueue(tchgeo_plan_t* p, uint64_t seed, uint32_t batch_base, cudaStream_t stream,
                                  std::vector<cudaEvent_t>* ev) {
  TCHGEO_REQUIRE(p != nullptr, "plan is NULL");
  TCHGEO_REQUIRE(!p->pending, "the previous enqueue of this plan has not been collected");
  p->a.stream = stream;
  p->a.seed = seed;
  p->a.batch_base = batch_base;
  tchgeo_status st = enqueue_step(&p->a, p->pl, seed, batch_base, stream, ev);
  if (st != TCHGEO_OK) return st;
  // the step's only read-back: length table + error word into pinned memory, then the event collect() waits for
  char* ws = (char*)p->a.workspace;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(p->h_state, ws + p->pl.off_state, p->n_state * 8, cudaMemcpyDeviceToHost, stream));
  p->h_state[p->n_state] = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(p->h_state + p->n_state, ws + p->pl.off_ctrl, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaEventRecord(p->done, stream));
  p->pending = true;
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_plan_enqueue(tchgeo_plan_t* p, uint64_t seed, uint32_t batch_base, tchgeo_stream stream) {
  return plan_enqueue(p, seed, batch_base, (cudaStream_t)stream, nullptr);
}

extern "C" tchgeo_status tchgeo_plan_collect(tchgeo_plan_t* p) {
  TCHGEO_REQUIRE(p != nullptr, "plan is NULL");
  TCHGEO_REQUIRE(p->pending, "nothing to collect: call tchgeo_plan_enqueue first");
  p->pending = false;
  TCHGEO_CUDA_CHECK(cudaEventSynchronize(p->done));
  decode_state(p->pl, p->h_state, p->samples_len.data(), p->edges_len.data(), p->layer_offsets.data(),
               p->pl.relabel ? p->nodes_len.data() : nullptr);
  return status_from_dev_err((uint32_t)p->h_state[p->n_state]);
}

extern "C" tchgeo_status tchgeo_plan_enqueue_timed(tchgeo_plan_t* p, uint64_t seed, uint32_t batch_base,
                                                   tchgeo_stream stream, float* launch_ms, int32_t launch_ms_cap,
                                                   int32_t* num_intervals) {
  TCHGEO_REQUIRE(p != nullptr && launch_ms != nullptr, "NULL pointer");
  const int n_iv = (int)p->pl.launches.size() + (p->pl.relabel ? 1 : 0);
  if (num_intervals) *num_intervals = n_iv;
  TCHGEO_REQUIRE(launch_ms_cap >= n_iv, "launch_ms too small: %d intervals", n_iv);
  EventList evs;
  tchgeo_status st = make_events(evs, (size_t)n_iv + 1);
  if (st != TCHGEO_OK) return st;
  st = plan_enqueue(p, seed, batch_base, (cudaStream_t)stream, &evs.ev);
  if (st != TCHGEO_OK) return st;
  st = tchgeo_plan_collect(p);
  read_intervals(evs, launch_ms);
  return st;
}

extern "C" tchgeo_status tchgeo_plan_results(const tchgeo_plan_t* p, const int64_t** samples_len, const int64_t** edges_len,
                                             const int64_t** layer_offsets, const int64_t** nodes_len) {
  TCHGEO_REQUIRE(p != nullptr, "plan is NULL");
  if (samples_len) *samples_len = p->samples_len.data();
  if (edges_len) *edges_len = p->edges_len.data();
  if (layer_offsets) *layer_offsets = p->layer_offsets.data();
  if (nodes_len) *nodes_len = p->pl.relabel ? p->nodes_len.data() : nullptr;
  return TCHGEO_OK;
}

extern "C" int32_t tchgeo_plan_num_launches(const tchgeo_plan_t* p) {
  if (!p) return 0;
  int n = (int)p->pl.launches.size() + p->pl.relabel_kernels;
  for (int t = 0; t < p->pl.T; ++t)
    if (p->a.seeds_per_batch[t] > 0) ++n;  // fill_i64_kernel
  return n;
}

extern "C" tchgeo_status tchgeo_neighbor_sampling_homogenous(
    const int64_t* col_ptrs, int64_t num_cols, const int64_t* row_indices, const int64_t* inputs, int64_t num_batches,
    int64_t seeds_per_batch, const int64_t* num_neighbors, int32_t num_hops, int32_t sampler_kind,
    const double* weights, uint64_t seed, uint32_t batch_base, int64_t* samples, int64_t samples_stride, int64_t* rows,
    int64_t* cols, int64_t* edge_index, int64_t edges_stride, int64_t* out_lens, int64_t* layer_offsets,
    void* workspace, size_t workspace_bytes, tchgeo_stream stream) {
  const int32_t zero = 0;
  tchgeo_sampling_args a;
  memset(&a, 0, sizeof(a));
  a.num_node_types = 1; a.num_rels = 1; a.num_hops = num_hops; a.sampler_kind = sampler_kind;
  a.rel_src = &zero; a.rel_dst = &zero;
  a.col_ptrs = &col_ptrs; a.num_cols = &num_cols; a.row_indices = &row_indices; a.weights = &weights;
  a.fanouts = num_neighbors; a.rel_active = nullptr;
  a.num_batches = num_batches; a.inputs = &inputs; a.seeds_per_batch = &seeds_per_batch;
  a.seed = seed; a.batch_base = batch_base;
  a.samples = &samples; a.samples_stride = &samples_stride;
  a.rows = &rows; a.cols = &cols; a.edge_index = &edge_index; a.edges_stride = &edges_stride;
  std::vector<int64_t> slen, elen;
  if (out_lens) {
    slen.resize((size_t)std::max<int64_t>(num_batches, 1));
    elen.resize((size_t)std::max<int64_t>(num_batches, 1));
    a.samples_len = slen.data();
    a.edges_len = elen.data();
  }
  a.layer_offsets = layer_offsets;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes; a.stream = stream;
  tchgeo_status st = tchgeo_neighbor_sampling(&a);
  if (out_lens && (st == TCHGEO_OK || st >= TCHGEO_ERR_CAPACITY))
    for (int64_t b = 0; b < num_batches; ++b) {
      out_lens[2 * b] = slen[(size_t)b];
      out_lens[2 * b + 1] = elen[(size_t)b];
    }
  return st;
}

extern "C" tchgeo_status tchgeo_compress_indices(const int64_t* src, int64_t n, int32_t* dst, int32_t* scratch,
                                                 tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n >= 0 && scratch != nullptr && (n == 0 || (src && dst)), "bad compress argument");
  if (n == 0) return TCHGEO_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4, stream));
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  compress_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(src, n, dst, (uint32_t*)scratch);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t herr = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(herr);
}
