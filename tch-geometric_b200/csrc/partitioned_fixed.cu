// Range-partitioned CSC (BASELINE config 5), device-only protocol ("fixed segments").
//
// Every rank owns a column range of the CSC and samples its OWN seed batches; per hop the frontier is exchanged with
// the owners and the sampled neighbours come back (the frontier exchange north_star asks for), but neither exchange
// involves the host: every (requester q, owner o) pair has a FIXED segment of `seg` rows in o's request buffer and in
// q's answer buffer (NVLink peer memory), so nobody needs the count matrix an all-to-all(v) is sized with.
//   scatter  (requester)  frontier -> request rows (id, batch << 32 | pos), grouped by owner in a local send buffer with
//                         one cursor per owner (warp-aggregated); remembers for every frontier node where its answer
//                         will arrive (slot map)
//   put      (requester)  group o goes to rows [me*seg, me*seg + cnt_o) of owner o's request buffer in whole lines, and
//                         cnt_o to o's count table                                       -- peer stores, then a barrier
//   serve    (owner)      samples every request it holds (counts read from its count table, on the device) with exactly
//                         the counters of the replicated path and stores compact int32 answer rows (k ids + k LOCAL
//                         csc positions, -1 padded) at the SAME row index of the requester's answer segment: a tile's
//                         rows are contiguous there, so the store is one linear run        -- peer stores, then a barrier
//   finish   (requester)  ONE pass in frontier order: answer row through the slot map, per-node counts, block scan,
//                         decoupled look-back per batch, then the tree layout of src/algo/neighbor_sampling.rs:210-218
//                         (samples, rows, cols, edge_index) and the new lengths.
// A step is 4 launches per hop and two barriers; the only host synchronisation is the read-back of the lengths at the end
// of the call.  A segment that would overflow (`seg` = slack x the mean load) raises TCHGEO_ERR_CAPACITY.
// The draws depend on (seed, batch, position, degree) only, so the result equals tchgeo_neighbor_sampling bit for bit.
// Integer gather/scatter work bounded by HBM and, for the two exchanges, by NVLink: no tensor cores.
#include <algorithm>

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int PF_THREADS = 256;       // scatter / put
constexpr int PF_MAX_WORLD = 64;
constexpr int SV_THREADS = 128;       // serve / finish: one request (frontier node) per thread of a tile
constexpr int SV_LIGHT_CHUNKS = 16;   // 8-step draw chunks above which a request is strided by the whole CTA
constexpr int SV_MAX_FANOUT = 64;     // answer rows are staged in shared memory: 128 x 2k words
constexpr uint32_t SLOT_IDX_BITS = 26;
constexpr uint32_t SLOT_NONE = 0xFFFFFFFFu;
constexpr uint64_t ST_FLAG_AGG = 1ull << 62;
constexpr uint64_t ST_FLAG_INCL = 2ull << 62;
constexpr uint64_t ST_VAL_MASK = (1ull << 62) - 1;

struct Frontier {
  const int64_t* samples;     // [B, samples_stride]
  int64_t samples_stride;
  const int64_t* fr_begin;    // [B] or NULL (= 0)
  const int64_t* fr_end;      // [B]
  int64_t B, capF;            // capF: upper bound of the per-batch frontier size (launch geometry)
};

__device__ __forceinline__ bool fr_node(const Frontier& f, int64_t b, int64_t j, int64_t& pos, int64_t& id) {
  if (b >= f.B || j >= f.capF) return false;
  const int64_t fb = f.fr_begin ? f.fr_begin[b] : 0;
  int64_t fe = f.fr_end[b];
  if (fe > f.samples_stride) fe = f.samples_stride;
  if (j >= fe - fb) return false;
  pos = fb + j;
  id = f.samples[b * f.samples_stride + pos];
  return true;
}

__device__ __forceinline__ int owner_of(int64_t id, int64_t cols_per_rank, int world) {
  int64_t o;
  if (id < 0) o = 0;
  else if (((uint64_t)id | (uint64_t)cols_per_rank) >> 32) o = id / cols_per_rank;
  else o = (int64_t)((uint32_t)id / (uint32_t)cols_per_rank);  // the common case: a 32-bit division
  if (o >= world) o = world - 1;  // out-of-range ids go to the edge ranks, whose serve kernel reports them
  return (int)o;
}

// ---- scatter: frontier -> local send buffer grouped by owner ---------------------------------------------------
struct ScatterParams {
  Frontier f;
  int64_t cols_per_rank, seg;
  int32_t world;
  uint32_t batch_base;
  unsigned long long* cursor;  // [world] rows taken in every owner's group (zeroed before the launch)
  int64_t* send;               // [world, seg, 2]
  uint32_t* slot_of;           // [B * capF] owner << 26 | row inside the segment
  uint32_t* err;
};

__global__ void __launch_bounds__(PF_THREADS) pf_scatter_kernel(const ScatterParams p) {
  __shared__ unsigned int hist[PF_MAX_WORLD];
  __shared__ unsigned long long base[PF_MAX_WORLD];
  if (threadIdx.x < PF_MAX_WORLD) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t b = blockIdx.y, j = (int64_t)blockIdx.x * PF_THREADS + threadIdx.x;
  int64_t pos = 0, id = 0;
  if (j == 0 && p.f.fr_end[b] - (p.f.fr_begin ? p.f.fr_begin[b] : 0) > p.f.capF) atomicOr(p.err, DEV_ERR_CAPACITY);
  const bool ok = fr_node(p.f, b, j, pos, id);
  const int o = ok ? owner_of(id, p.cols_per_rank, p.world) : 0;
  // warp-aggregated histogram: lanes with the same owner elect a leader that adds the group's size once
  const unsigned lane = threadIdx.x & 31u;
  const unsigned m = __match_any_sync(0xffffffffu, ok ? o : -1);
  const int leader = __ffs(m) - 1;
  unsigned int r0 = 0;
  if (ok && (int)lane == leader) r0 = atomicAdd(&hist[o], (unsigned int)__popc(m));
  r0 = __shfl_sync(0xffffffffu, r0, leader);
  const unsigned int rank = r0 + (unsigned int)__popc(m & ((1u << lane) - 1u));
  __syncthreads();
  if (threadIdx.x < p.world) {
    const unsigned int h = hist[threadIdx.x];
    base[threadIdx.x] = h ? atomicAdd(p.cursor + threadIdx.x, (unsigned long long)h) : 0ull;
  }
  __syncthreads();
  if (ok) {
    const unsigned long long q = base[o] + rank;
    uint32_t slot = SLOT_NONE;
    if (q < (unsigned long long)p.seg) {
      const uint64_t meta = ((uint64_t)(p.batch_base + (uint32_t)b) << 32) | (uint64_t)(uint32_t)pos;
      *reinterpret_cast<longlong2*>(p.send + 2 * ((int64_t)o * p.seg + (int64_t)q)) =
          make_longlong2((long long)id, (long long)meta);
      slot = ((uint32_t)o << SLOT_IDX_BITS) | (uint32_t)q;
    } else {
      atomicOr(p.err, DEV_ERR_CAPACITY);  // this owner's segment is full: the plan needs a larger slack
    }
    p.slot_of[b * p.f.capF + j] = slot;
  }
}

// ---- put: every owner's group travels to that owner's request buffer in whole lines ---------------------------------
struct PutParams {
  const int64_t* send;                 // [world, seg, 2]
  const unsigned long long* cursor;    // [world] rows per group (may exceed seg after an overflow: clamped)
  int64_t seg;
  int32_t world, me;
  int64_t* peer_req[PF_MAX_WORLD];     // owner o's request buffer [world, seg, 2]
  int64_t* peer_cnt[PF_MAX_WORLD];     // owner o's count table [world]
};

__global__ void __launch_bounds__(PF_THREADS) pf_put_kernel(const PutParams p) {
  // peers are visited in a rank-dependent rotation: at any moment the ranks then store into DIFFERENT peers (a
  // permutation schedule).  With the same order on every rank all of them hit one receiver at a time (incast), which
  // held the 8-GPU exchange at ~200 GB/s per rank.
  const int o = (int)((blockIdx.y + (unsigned)p.me + 1u) % (unsigned)p.world);
  unsigned long long n = p.cursor[o];
  if (n > (unsigned long long)p.seg) n = (unsigned long long)p.seg;
  if (blockIdx.x == 0 && threadIdx.x == 0) p.peer_cnt[o][p.me] = (int64_t)n;
  const longlong2* src = reinterpret_cast<const longlong2*>(p.send) + (int64_t)o * p.seg;
  longlong2* dst = reinterpret_cast<longlong2*>(p.peer_req[o]) + (int64_t)p.me * p.seg;
  const int64_t q0 = (int64_t)blockIdx.x * (PF_THREADS * 4);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t q = q0 + u * PF_THREADS + threadIdx.x;
    if ((unsigned long long)q < n) dst[q] = src[q];
  }
}

// ---- serve: the owner samples the requests of every requester's segment ------------------------------------------
struct ServeParams {
  const int64_t* ptrs;        // local colptr [ncols+1], rebased to the local indices arrays
  const int64_t* indices;
  const int32_t* indices32;   // optional int32 replica of the local indices
  const double* weights;
  const int64_t* req_in;      // [world, seg, 2]
  const int64_t* cnt_in;      // [world]
  int64_t col_begin, ncols, nnz, seg;
  int32_t fanout, world, me;
  uint32_t key0, key1, rel;
  uint32_t rk[20];            // Philox round keys (constant bank)
  int32_t* peer_ans[PF_MAX_WORLD];  // requester q's answer buffer [world, seg, 2k]
  uint32_t* err;
};

__device__ __forceinline__ Philox4 philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return Philox4{c0, c1, c2, c3};
}

// steps step0 .. step0+3 of the serial reservoir (sampling.rs:17-23): a hit (j < k) overwrites slot j, the LAST hit of a
// slot wins -> max of the step index
// (`slots_sa` is a 32-bit shared-state-space address; the update is a predicated red.shared.max: no branch)
__device__ __forceinline__ void reservoir_block(const Philox4& r, uint32_t step0, uint32_t deg, uint32_t k, uint32_t slots_sa) {
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

template <int KIND, bool I32>
__global__ void __launch_bounds__(SV_THREADS, 10) pf_serve_kernel(const ServeParams p) {
  constexpr int NT = SV_THREADS, NW = NT / 32;
  __shared__ int64_t s_start[NT];
  __shared__ uint32_t s_deg[NT], s_pos[NT], s_batch[NT], s_choff[NT];
  __shared__ uint8_t s_chown[SV_LIGHT_CHUNKS * NT];
  __shared__ uint8_t s_heavy[NT];
  __shared__ uint32_t s_wtot[NW];
  __shared__ uint32_t s_nheavy;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t k = (uint32_t)p.fanout;
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(dyn_smem);           // [NT * k] chosen position inside the column
  int32_t* s_ans = reinterpret_cast<int32_t*>(s_slot + (size_t)NT * k);  // [NT * 2k] the tile's answer rows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // requester whose segment this CTA serves; rotated by the rank so that the ranks' answer stores go to different
  // peers at any moment (see pf_put_kernel)
  const int q = (int)((blockIdx.y + (unsigned)p.me + 1u) % (unsigned)p.world);
  int64_t n_req = p.cnt_in[q];
  if (n_req > p.seg) n_req = p.seg;
  const int64_t r0 = (int64_t)blockIdx.x * NT;
  if (r0 >= n_req) return;
  const int nn = (int)min((int64_t)NT, n_req - r0);
  if (tid == 0) s_nheavy = 0u;

  // ---- A: request rows (coalesced 16-byte loads), colptr pairs, degrees ------------------------------------------
  uint32_t deg = 0, nch = 0;
  int64_t start = 0;
  bool heavy = false;
  if (tid < nn) {
    const longlong2 row = *(reinterpret_cast<const longlong2*>(p.req_in) + (int64_t)q * p.seg + r0 + tid);
    const int64_t w = row.x - p.col_begin;
    s_pos[tid] = (uint32_t)row.y;
    s_batch[tid] = (uint32_t)((uint64_t)row.y >> 32);
    if (w < 0 || w >= p.ncols) {
      atomicOr(p.err, DEV_ERR_INDEX);  // reference: slice index panic (quirk Q10)
    } else {
      const uint64_t keep = l2_policy_evict_last();
      start = ld_gather64_keep_i64(p.ptrs + w, keep);
      const int64_t end = ld_gather64_keep_i64(p.ptrs + w + 1, keep);
      const int64_t d = end - start;
      if (d < 0 || d > 0x7fffffffll || start < 0 || end > p.nnz) atomicOr(p.err, DEV_ERR_INDEX);
      else deg = (uint32_t)d;
    }
    if (KIND != TCHGEO_SAMPLER_UNIFORM_REPLACE && k == 0 && deg > 0) atomicOr(p.err, DEV_ERR_PANIC);  // gen_range(0..0)
    if (KIND == TCHGEO_SAMPLER_UNIFORM && deg > k) {
      nch = (deg - k + 7u) >> 3;
      heavy = nch > (uint32_t)SV_LIGHT_CHUNKS;
    }
  }
  s_start[tid] = start;
  s_deg[tid] = deg;
  const uint32_t light = heavy ? 0u : nch;
  uint32_t incl = light;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t choff = incl - light, Q = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const uint32_t v = s_wtot[w];
    if (w < warp) choff += v;
    Q += v;
  }
  s_choff[tid] = choff;
  if (tid < nn) {
    // slot s starts as "item s" and is raised to the step index of every hit (steps >= k > s): after the draws it is the
    // position of the chosen neighbour (last hit, else item s; sampling.rs:17-23)
    for (uint32_t s = 0; s < k; ++s) s_slot[tid * k + s] = s;
    for (uint32_t c = 0; c < light; ++c) s_chown[choff + c] = (uint8_t)tid;
    if (heavy) s_heavy[atomicAdd(&s_nheavy, 1u)] = (uint8_t)tid;
  }
  __syncthreads();

  // ---- C: sampling decisions in shared memory (same Philox counters as hop_kernel) ------------------------------
  if (KIND == TCHGEO_SAMPLER_UNIFORM) {
    const uint32_t tag = TAG_RESERVOIR | (p.rel << 8);
    const uint32_t slot_sa = (uint32_t)__cvta_generic_to_shared(s_slot);
#pragma unroll 1
    for (uint32_t qi = tid; qi < Q; qi += NT) {  // (request, 8-step chunk) work items: Philox blocks 2c and 2c+1
      const uint32_t n = s_chown[qi];
      const uint32_t c = qi - s_choff[n];
      const uint32_t dn = s_deg[n], step0 = k + 8u * c;
      const uint32_t sa = slot_sa + 4u * n * k;
      reservoir_block(philox_rk(s_pos[n], 2u * c, s_batch[n], tag, p.rk), step0, dn, k, sa);
      if (step0 + 4u < dn) reservoir_block(philox_rk(s_pos[n], 2u * c + 1u, s_batch[n], tag, p.rk), step0 + 4u, dn, k, sa);
    }
    const uint32_t nheavy = s_nheavy;
    for (uint32_t h = 0; h < nheavy; ++h) {   // hubs: the whole CTA strides over one request's 4-step blocks
      const uint32_t n = s_heavy[h];
      const uint32_t dn = s_deg[n], nb = (dn - k + 3u) >> 2;
#pragma unroll 1
      for (uint32_t c = tid; c < nb; c += NT)
        reservoir_block(philox_rk(s_pos[n], c, s_batch[n], tag, p.rk), k + 4u * c, dn, k, slot_sa + 4u * n * k);
    }
  } else if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
    const uint32_t rtag = TAG_REPLACE | (p.rel << 8);
    const uint32_t bpn = (k + 3u) >> 2;
#pragma unroll 1
    for (uint32_t qi = tid; qi < (uint32_t)nn * bpn; qi += NT) {  // k iid picks (sampling.rs:57-69), one Philox block per four
      const uint32_t n = qi / bpn, c = qi - n * bpn;
      const uint32_t dn = s_deg[n];
      if (dn == 0) continue;
      const Philox4 r = philox_rk(s_pos[n], c, s_batch[n], rtag, p.rk);
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u)
        if (4u * c + u < k) s_slot[n * k + 4u * c + u] = __umulhi(pick4(r, u), dn);
    }
  } else {
    for (int n = warp; n < nn; n += NW) {  // weighted: warp per request, f64 shuffle scan of the weights (sampling.rs:28-55)
      const uint32_t dn = s_deg[n];
      if (dn <= k) continue;
      const double* wp = p.weights + s_start[n];
      double carry = 0.0;
      for (uint32_t base = 0; base < dn; base += 32) {
        const uint32_t item = base + lane;
        const double w = item < dn ? __ldg(wp + item) : 0.0;
        double iw = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, iw, o);
          if (lane >= o) iw += up;
        }
        const double w_sum = carry + iw;
        if (item >= k && item < dn) {
          if (!(w_sum > 0.0)) {
            atomicOr(p.err, DEV_ERR_PANIC);
          } else {
            const Philox4 r = philox_rk(s_pos[n], item, s_batch[n], TAG_WEIGHTED | (p.rel << 8), p.rk);
            const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
            const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
            if (__dmul_rn(u, w_sum) < w) atomicMax(s_slot + n * k + __umulhi(r.z, k), item);
          }
        }
        carry = __shfl_sync(0xffffffffu, w_sum, 31);
      }
    }
  }
  __syncthreads();

  // ---- D: gathers, answer rows staged in shared memory, one linear run of stores into the requester's segment ----
  const uint32_t w2 = 2u * k;
  constexpr uint32_t U = 5;  // gathers in flight per thread (the whole tile in one round when fanout <= 5)
  const uint32_t total = (uint32_t)nn * k;
#pragma unroll 1
  for (uint32_t e0 = tid; e0 < total; e0 += NT * U) {
    uint32_t at[U];
    int32_t id[U], lp32[U];
#pragma unroll
    for (uint32_t u = 0; u < U; ++u) {
      const uint32_t e = e0 + u * NT;
      id[u] = -1; lp32[u] = -1; at[u] = 0xffffffffu;
      if (e < total) {
        const uint32_t n = e / k, s = e - n * k;
        const uint32_t dn = s_deg[n];
        const uint32_t cnt = KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE ? (dn > 0 ? k : 0u) : min(dn, k);
        at[u] = n * w2 + s;
        if (s < cnt) {
          const int64_t lp = s_start[n] + s_slot[e];
          id[u] = I32 ? ld_gather64_i32(p.indices32 + lp) : (int32_t)ld_gather64_i64(p.indices + lp);
          lp32[u] = (int32_t)lp;
        }
      }
    }
#pragma unroll
    for (uint32_t u = 0; u < U; ++u)
      if (at[u] != 0xffffffffu) {
        s_ans[at[u]] = id[u];
        s_ans[at[u] + k] = lp32[u];
      }
  }
  __syncthreads();
  // the tile's rows are one contiguous run in the requester's segment: 16-byte stores when the run is 16-byte aligned
  // (NVLink peer stores bypass the L2: every store instruction is a fabric transaction, so wider is cheaper)
  int32_t* dst = p.peer_ans[q] + ((int64_t)p.me * p.seg + r0) * w2;
  const uint32_t words = (uint32_t)nn * w2;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(s_ans);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = tid; i < (words >> 2); i += NT) d4[i] = s4[i];
    for (uint32_t i = (words & ~3u) + tid; i < words; i += NT) dst[i] = s_ans[i];
  } else {
    for (uint32_t i = tid; i < words; i += NT) dst[i] = s_ans[i];
  }
}

// ---- finish: answers -> tree layout, one pass in frontier order ------------------------------------------------------
struct FinishParams {
  Frontier f;
  const int32_t* ans_in;      // [world, seg, 2k]
  const uint32_t* slot_of;    // [B * capF]
  const int64_t* edge_base;   // [world] CSC entries owned by lower ranks
  int64_t seg;
  int32_t k, tiles_per_batch;
  uint32_t total_tiles;
  const int64_t* node_len_in;  // [B]
  const int64_t* edge_len_in;
  int64_t* node_len_out;
  int64_t* edge_len_out;
  int64_t* samples;            // [B, samples_stride] (also holds the frontier: f.samples)
  int64_t* rows; int64_t* cols; int64_t* eidx; int64_t edges_stride;
  uint64_t* status;            // [B * tiles_per_batch] look-back words, zero-initialised
  uint32_t* ticket;
  uint32_t* err;
};

__global__ void __launch_bounds__(SV_THREADS, 8) pf_finish_kernel(const FinishParams p) {
  constexpr int NT = SV_THREADS, NW = NT / 32;
  __shared__ uint32_t s_off[NT + 1];
  __shared__ uint32_t s_owner_rank[NT];
  __shared__ uint32_t s_wtot[NW];
  __shared__ uint32_t s_ticket;
  __shared__ int64_t s_excl;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t k = (uint32_t)p.k, w2 = 2u * k;
  int32_t* s_rows = reinterpret_cast<int32_t*>(dyn_smem);                 // [NT * 2k] the tile's answer rows
  uint8_t* s_owner = reinterpret_cast<uint8_t*>(s_rows + (size_t)NT * w2);  // [NT * k] output edge -> node of the tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles are handed out in start order and tile-major, so the per-batch look-back chains advance independently and
  // every tile a look-back waits for belongs to a CTA that is already running
  if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const uint32_t ticket = s_ticket;
  if (ticket >= p.total_tiles) return;
  const int t = (int)(ticket / (uint32_t)p.f.B);
  const int64_t b = (int64_t)(ticket - (uint32_t)t * (uint32_t)p.f.B);
  const int64_t fb = p.f.fr_begin ? p.f.fr_begin[b] : 0;
  int64_t fe = p.f.fr_end[b];
  if (fe > p.f.samples_stride) fe = p.f.samples_stride;
  const int64_t F = fe > fb ? min(fe - fb, p.f.capF) : 0;
  const int64_t node0 = (int64_t)t * NT;
  const int nn = F > node0 ? (int)min((int64_t)NT, F - node0) : 0;
  const bool is_last = (nn > 0 && node0 + nn == F) || (F == 0 && t == 0);
  if (nn == 0 && !is_last) return;

  // ---- the node's answer row (through the slot map) and its number of answers (valid slots form a prefix) --------
  uint32_t cnt = 0, slot = SLOT_NONE;
  if (tid < nn) {
    slot = p.slot_of[b * p.f.capF + node0 + tid];
    if (slot != SLOT_NONE) {
      const int64_t row = (int64_t)(slot >> SLOT_IDX_BITS) * p.seg + (int64_t)(slot & ((1u << SLOT_IDX_BITS) - 1u));
      const int32_t* a = p.ans_in + row * w2;
      int32_t* d = s_rows + (size_t)tid * w2;
      if ((w2 & 1u) == 0 && ((row * w2) & 1) == 0) {   // rows of 2k words are 8-byte aligned: two words per load
        for (uint32_t s = 0; s < w2; s += 2) {
          const int2 v = __ldcs(reinterpret_cast<const int2*>(a + s));
          d[s] = v.x; d[s + 1] = v.y;
        }
      } else {
        for (uint32_t s = 0; s < w2; ++s) d[s] = __ldcs(a + s);
      }
      for (uint32_t s = 0; s < k; ++s) cnt += d[k + s] >= 0 ? 1u : 0u;
    }
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t off = incl - cnt, total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const uint32_t v = s_wtot[w];
    if (w < warp) off += v;
    total += v;
  }
  uint64_t* my_status = p.status + (size_t)b * p.tiles_per_batch;
  if (tid == 0) st_relaxed_u64(my_status + t, (t == 0 ? ST_FLAG_INCL : ST_FLAG_AGG) | (uint64_t)total);
  s_off[tid] = off;
  s_owner_rank[tid] = slot == SLOT_NONE ? 0u : (slot >> SLOT_IDX_BITS);
  for (uint32_t s = 0; s < cnt; ++s) s_owner[off + s] = (uint8_t)tid;

  // ---- decoupled look-back over this batch's earlier tiles (warp 0) ---------------------------------------------
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
        if (first_inval < first_incl) {
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
  __syncthreads();
  const int64_t excl = s_excl;
  const int64_t e_base = p.edge_len_in[b] + excl, s_base = p.node_len_in[b] + excl;
  if (is_last && tid == 0) {
    p.edge_len_out[b] = e_base + total;
    p.node_len_out[b] = s_base + total;
  }
  if (e_base + total > p.edges_stride || s_base + total > p.f.samples_stride) {
    if (tid == 0) atomicOr(p.err, DEV_ERR_CAPACITY);
    return;
  }
  // ---- the tree layout of neighbor_sampling.rs:210-218: one thread per output edge, coalesced streaming stores ----
  int64_t* const ps = p.samples + b * p.f.samples_stride + s_base;
  int64_t* const pr = p.rows + b * p.edges_stride + e_base;
  int64_t* const pc = p.cols + b * p.edges_stride + e_base;
  int64_t* const pe = p.eidx + b * p.edges_stride + e_base;
  const int64_t col0 = fb + node0;
#pragma unroll 1
  for (uint32_t e = tid; e < total; e += NT) {
    const uint32_t n = s_owner[e], s = e - s_off[n];
    const int32_t* a = s_rows + (size_t)n * w2;
    st_cs_i64(ps + e, (int64_t)a[s]);
    st_cs_i64(pr + e, s_base + e);                        // index of the appended node
    st_cs_i64(pc + e, col0 + n);                          // index of the frontier node
    st_cs_i64(pe + e, p.edge_base[s_owner_rank[n]] + (int64_t)a[k + s]);  // global CSC position
  }
}

struct FinishWs {
  size_t off_ticket, off_status, total;
};
bool finish_ws(int64_t num_batches, int64_t frontier_cap, FinishWs& w) {
  if (num_batches <= 0 || frontier_cap < 0) return false;
  const int64_t tpb = std::max<int64_t>(1, (frontier_cap + SV_THREADS - 1) / SV_THREADS);
  if (tpb * num_batches >= ((int64_t)1 << 31)) return false;
  w.off_ticket = 0;
  w.off_status = 256;
  w.total = 256 + ((size_t)tpb * num_batches * 8 + 255) / 256 * 256;
  return true;
}

template <typename F>
tchgeo_status configure_smem(F kernel, size_t smem) {
  if (smem > 48 * 1024) TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return TCHGEO_OK;
}

template <int KIND>
tchgeo_status launch_serve(const ServeParams& sp, dim3 grid, size_t smem, cudaStream_t stream) {
  tchgeo_status st;
  if (sp.indices32) {
    st = configure_smem(pf_serve_kernel<KIND, true>, smem);
    if (st != TCHGEO_OK) return st;
    pf_serve_kernel<KIND, true><<<grid, SV_THREADS, smem, stream>>>(sp);
  } else {
    st = configure_smem(pf_serve_kernel<KIND, false>, smem);
    if (st != TCHGEO_OK) return st;
    pf_serve_kernel<KIND, false><<<grid, SV_THREADS, smem, stream>>>(sp);
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" size_t tchgeo_partf_workspace_bytes(int64_t num_batches, int64_t frontier_cap) {
  FinishWs w;
  return finish_ws(num_batches, frontier_cap, w) ? w.total : 0;
}

extern "C" tchgeo_status tchgeo_partf_scatter(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                              const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                              int64_t cols_per_rank, int32_t world, int32_t me, uint32_t batch_base,
                                              int64_t seg_rows, int64_t* send, int64_t* cursor, uint32_t* slot_of,
                                              void* const* peer_req, void* const* peer_cnt, int32_t* err_word,
                                              tchgeo_stream stream_) {
  TCHGEO_REQUIRE(world >= 1 && world <= PF_MAX_WORLD && me >= 0 && me < world, "bad world / rank");
  TCHGEO_REQUIRE(num_batches >= 0 && num_batches <= 65535 && frontier_cap >= 0 && cols_per_rank >= 1 && samples_stride >= 0,
                 "bad argument");
  TCHGEO_REQUIRE(seg_rows >= 1 && seg_rows < ((int64_t)1 << SLOT_IDX_BITS), "segment size must be in [1, 2^26)");
  TCHGEO_REQUIRE(send && cursor && slot_of && peer_req && peer_cnt && err_word, "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(cursor, 0, (size_t)world * 8, stream));
  PutParams pp;
  pp.send = send; pp.cursor = (const unsigned long long*)cursor; pp.seg = seg_rows; pp.world = world; pp.me = me;
  for (int o = 0; o < world; ++o) {
    TCHGEO_REQUIRE(peer_req[o] && peer_cnt[o], "peer table entry %d is NULL", o);
    pp.peer_req[o] = (int64_t*)peer_req[o];
    pp.peer_cnt[o] = (int64_t*)peer_cnt[o];
  }
  if (num_batches > 0 && frontier_cap > 0) {
    TCHGEO_REQUIRE(samples && fr_end, "NULL pointer");
    ScatterParams sp;
    sp.f.samples = samples; sp.f.samples_stride = samples_stride; sp.f.fr_begin = fr_begin; sp.f.fr_end = fr_end;
    sp.f.B = num_batches; sp.f.capF = frontier_cap;
    sp.cols_per_rank = cols_per_rank; sp.seg = seg_rows; sp.world = world; sp.batch_base = batch_base;
    sp.cursor = (unsigned long long*)cursor; sp.send = send; sp.slot_of = slot_of; sp.err = (uint32_t*)err_word;
    const int64_t gx = (frontier_cap + PF_THREADS - 1) / PF_THREADS;
    TCHGEO_REQUIRE(gx < ((int64_t)1 << 31), "frontier too large for one launch");
    pf_scatter_kernel<<<dim3((unsigned)gx, (unsigned)num_batches), PF_THREADS, 0, stream>>>(sp);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  // the groups cannot hold more rows than the frontier has: bound the put grid by both
  const int64_t rows_max = std::max<int64_t>(1, std::min<int64_t>(seg_rows, num_batches * frontier_cap));
  const int64_t gx = (rows_max + PF_THREADS * 4 - 1) / (PF_THREADS * 4);
  pf_put_kernel<<<dim3((unsigned)gx, (unsigned)world), PF_THREADS, 0, stream>>>(pp);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_partf_serve(const int64_t* ptrs_local, const int64_t* indices_local,
                                            const int32_t* indices32_local, const double* weights_local, int64_t col_begin,
                                            int64_t ncols_local, int64_t nnz_local, const int64_t* req_in,
                                            const int64_t* cnt_in, int64_t seg_rows, int64_t max_requests, int64_t fanout,
                                            int32_t sampler_kind, uint64_t seed, uint32_t rel, int32_t world, int32_t me,
                                            void* const* peer_ans, int32_t* err_word, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(world >= 1 && world <= PF_MAX_WORLD && me >= 0 && me < world, "bad world / rank");
  TCHGEO_REQUIRE(fanout >= 0 && fanout <= SV_MAX_FANOUT, "fanout must be in [0, %d] on the partitioned path", SV_MAX_FANOUT);
  TCHGEO_REQUIRE(sampler_kind >= 0 && sampler_kind <= 2 && ncols_local >= 0 && seg_rows >= 1, "bad serve argument");
  TCHGEO_REQUIRE(nnz_local >= 0 && nnz_local < ((int64_t)1 << 31), "a rank's share of the CSC must stay below 2^31 entries");
  TCHGEO_REQUIRE(sampler_kind != TCHGEO_SAMPLER_WEIGHTED || weights_local != nullptr, "weighted serve without weights");
  TCHGEO_REQUIRE(ptrs_local && req_in && cnt_in && peer_ans && err_word, "NULL pointer");
  if (fanout == 0 && sampler_kind == TCHGEO_SAMPLER_UNIFORM_REPLACE) return TCHGEO_OK;
  ServeParams sp;
  sp.ptrs = ptrs_local; sp.indices = indices_local; sp.indices32 = indices32_local; sp.weights = weights_local;
  sp.req_in = req_in; sp.cnt_in = cnt_in;
  sp.col_begin = col_begin; sp.ncols = ncols_local; sp.nnz = nnz_local; sp.seg = seg_rows;
  sp.fanout = (int32_t)fanout; sp.world = world; sp.me = me;
  sp.key0 = (uint32_t)seed; sp.key1 = (uint32_t)(seed >> 32); sp.rel = rel;
  for (uint32_t r = 0; r < 10; ++r) {
    sp.rk[2 * r] = sp.key0 + r * 0x9E3779B9u;
    sp.rk[2 * r + 1] = sp.key1 + r * 0xBB67AE85u;
  }
  for (int q = 0; q < world; ++q) {
    TCHGEO_REQUIRE(peer_ans[q] != nullptr, "peer answer table entry %d is NULL", q);
    sp.peer_ans[q] = (int32_t*)peer_ans[q];
  }
  sp.err = (uint32_t*)err_word;
  const int64_t rows_max = std::max<int64_t>(1, std::min<int64_t>(seg_rows, max_requests > 0 ? max_requests : seg_rows));
  const int64_t gx = (rows_max + SV_THREADS - 1) / SV_THREADS;
  TCHGEO_REQUIRE(gx < ((int64_t)1 << 31), "too many requests for one launch");
  const dim3 grid((unsigned)gx, (unsigned)world);
  const size_t smem = (size_t)SV_THREADS * std::max<int64_t>(fanout, 1) * 12 + 16;
  cudaStream_t stream = (cudaStream_t)stream_;
  switch (sampler_kind) {
    case TCHGEO_SAMPLER_UNIFORM: return launch_serve<TCHGEO_SAMPLER_UNIFORM>(sp, grid, smem, stream);
    case TCHGEO_SAMPLER_UNIFORM_REPLACE: return launch_serve<TCHGEO_SAMPLER_UNIFORM_REPLACE>(sp, grid, smem, stream);
    default: return launch_serve<TCHGEO_SAMPLER_WEIGHTED>(sp, grid, smem, stream);
  }
}

extern "C" tchgeo_status tchgeo_partf_finish(const int32_t* ans_in, const uint32_t* slot_of, int64_t seg_rows, int64_t fanout,
                                             const int64_t* owner_edge_base, int32_t world, const int64_t* fr_begin,
                                             const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                             const int64_t* node_len_in, const int64_t* edge_len_in, int64_t* node_len_out,
                                             int64_t* edge_len_out, int64_t* samples, int64_t samples_stride, int64_t* rows,
                                             int64_t* cols, int64_t* edge_index, int64_t edges_stride, int32_t* err_word,
                                             void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(fanout >= 0 && fanout <= SV_MAX_FANOUT && num_batches >= 0 && frontier_cap >= 0 && seg_rows >= 1, "bad argument");
  TCHGEO_REQUIRE(world >= 1 && world <= PF_MAX_WORLD, "bad world");
  TCHGEO_REQUIRE(node_len_in && edge_len_in && node_len_out && edge_len_out && err_word && owner_edge_base && fr_end,
                 "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_batches == 0) return TCHGEO_OK;
  if (frontier_cap == 0 || fanout == 0) {  // nothing can be appended: the lengths carry over
    TCHGEO_CUDA_CHECK(cudaMemcpyAsync(node_len_out, node_len_in, (size_t)num_batches * 8, cudaMemcpyDeviceToDevice, stream));
    TCHGEO_CUDA_CHECK(cudaMemcpyAsync(edge_len_out, edge_len_in, (size_t)num_batches * 8, cudaMemcpyDeviceToDevice, stream));
    return TCHGEO_OK;
  }
  FinishWs W;
  TCHGEO_REQUIRE(finish_ws(num_batches, frontier_cap, W), "frontier too large for one call");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= W.total, "workspace too small: need %zu bytes", W.total);
  TCHGEO_REQUIRE(ans_in && slot_of && samples && rows && cols && edge_index, "NULL pointer");
  TCHGEO_REQUIRE(samples_stride < ((int64_t)1 << 32), "samples_stride must be < 2^32");
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(workspace, 0, W.total, stream));
  FinishParams p;
  p.f.samples = samples; p.f.samples_stride = samples_stride; p.f.fr_begin = fr_begin; p.f.fr_end = fr_end;
  p.f.B = num_batches; p.f.capF = frontier_cap;
  p.ans_in = ans_in; p.slot_of = slot_of; p.edge_base = owner_edge_base; p.seg = seg_rows; p.k = (int32_t)fanout;
  p.tiles_per_batch = (int32_t)((frontier_cap + SV_THREADS - 1) / SV_THREADS);
  p.total_tiles = (uint32_t)((int64_t)p.tiles_per_batch * num_batches);
  p.node_len_in = node_len_in; p.edge_len_in = edge_len_in; p.node_len_out = node_len_out; p.edge_len_out = edge_len_out;
  p.samples = samples; p.rows = rows; p.cols = cols; p.eidx = edge_index; p.edges_stride = edges_stride;
  p.status = (uint64_t*)((char*)workspace + W.off_status);
  p.ticket = (uint32_t*)((char*)workspace + W.off_ticket);
  p.err = (uint32_t*)err_word;
  const size_t smem = (size_t)SV_THREADS * fanout * 9 + 16;
  tchgeo_status st = configure_smem(pf_finish_kernel, smem);
  if (st != TCHGEO_OK) return st;
  pf_finish_kernel<<<p.total_tiles, SV_THREADS, smem, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}
