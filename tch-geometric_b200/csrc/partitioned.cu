// Owner-side kernel of the range-partitioned CSC path (BASELINE config 5: graph larger than one GPU's
// share, columns split over ranks, hop frontiers exchanged with NCCL all-to-all).
//
// A rank receives REQUESTS (node id, batch index, position of the node in its batch's samples vector)
// for the columns it owns and answers each with the sampled neighbours: fanout slots per request,
// (neighbour id, GLOBAL csc position) or -1 for unused slots.  The draws use exactly the counters of the
// replicated path (DESIGN.md "RNG contract": they depend on seed, batch, position and the degree only),
// so the requester, which lays the answers out in frontier order, reproduces the single-GPU result bit
// for bit no matter how the graph is partitioned.  Nothing here depends on the reference beyond the
// samplers of src/utils/sampling.rs:6-69 restated in neighbor_sampling.cu.
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int SV_THREADS = 128;
constexpr int SV_LIGHT_MAX = 32;
constexpr int SV_MAX_TILE_SLOTS = 8192;
constexpr int SV_MAX_WORLD = 64;

struct ServeParams {
  const int64_t* ptrs;      // local colptr [ncols+1], rebased to the local indices array
  const int64_t* indices;   // local row_indices
  const double* weights;    // local weights or NULL
  const int64_t* req_ids;
  const int64_t* req_meta;  // (batch << 32) | pos
  int64_t req_stride;       // elements between consecutive requests (1: two arrays, 2: interleaved (id, meta) rows)
  int64_t* out_ids;
  int64_t* out_ptrs;
  int64_t out_stride;       // elements between consecutive answer rows (fanout, or 2*fanout for [n][ids | ptrs] rows)
  int32_t* out32;           // compact rows instead: [n][fanout ids | fanout LOCAL csc positions] as int32, -1 padded
  uint32_t* err;
  int64_t col_begin, ncols, edge_base, n;
  int32_t fanout, tile_reqs;
  uint32_t key0, key1, rel;
  // peer mode (tchgeo_serve_requests_rows_peer): the requests are grouped by requesting rank; request i of rank q's
  // segment [peer_seg[q], peer_seg[q+1]) is answered into rank q's OWN answer buffer (NVLink peer memory) at row
  // peer_row0[q] + (i - peer_seg[q]), i.e. where the answer all-to-all would have put it
  int32_t peer_world;                    // 0: answers go to out32 / out_ids
  int32_t* peer_base[SV_MAX_WORLD];
  int64_t peer_seg[SV_MAX_WORLD + 1];
  int64_t peer_row0[SV_MAX_WORLD];
};

__device__ __forceinline__ void sv_block(const Philox4& r, uint32_t step0, uint32_t deg, uint32_t k, uint32_t* slots) {
#pragma unroll
  for (uint32_t u = 0; u < 4; ++u) {
    const uint32_t step = step0 + u;
    const uint32_t j = __umulhi(pick4(r, u), step);
    if ((step < deg) & (j < k)) atomicMax(slots + j, step);
  }
}

template <int KIND>
__global__ void __launch_bounds__(SV_THREADS) serve_kernel(const ServeParams p) {
  using BlockScan = cub::BlockScan<uint32_t, SV_THREADS>;
  __shared__ typename BlockScan::TempStorage scan_tmp;
  __shared__ int64_t s_start[SV_THREADS];
  __shared__ uint32_t s_deg[SV_THREADS], s_pos[SV_THREADS], s_batch[SV_THREADS], s_choff[SV_THREADS];
  __shared__ uint8_t s_chown[SV_LIGHT_MAX * SV_THREADS];
  __shared__ uint8_t s_heavy[SV_THREADS];
  __shared__ uint32_t s_nheavy;
  __shared__ int32_t* s_row[SV_THREADS];  // compact answers: where request n's row goes (local buffer or a peer's)
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(dyn_smem);  // [tile_reqs * fanout]
  int32_t* s_ans = reinterpret_cast<int32_t*>(s_slot + (size_t)p.tile_reqs * p.fanout);  // [tile_reqs * 2 * fanout]

  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t k = (uint32_t)p.fanout;
  const int64_t r0 = (int64_t)blockIdx.x * p.tile_reqs;
  const int nn = (int)min((int64_t)p.tile_reqs, p.n - r0);
  if (tid == 0) s_nheavy = 0u;

  uint32_t deg = 0, nblocks = 0;
  int64_t start = 0;
  bool heavy = false;
  if (tid < nn && p.out32) s_row[tid] = p.out32 + (r0 + tid) * 2 * (int64_t)k;
  if (tid < nn && p.peer_world > 0) {
    const int64_t i = r0 + tid;
    int q = 0;
    while (q + 1 < p.peer_world && i >= p.peer_seg[q + 1]) ++q;
    s_row[tid] = p.peer_base[q] + (p.peer_row0[q] + (i - p.peer_seg[q])) * 2 * (int64_t)k;
  }
  if (tid < nn) {
    const int64_t w = p.req_ids[(r0 + tid) * p.req_stride] - p.col_begin;
    const int64_t meta = p.req_meta[(r0 + tid) * p.req_stride];
    s_pos[tid] = (uint32_t)meta;
    s_batch[tid] = (uint32_t)((uint64_t)meta >> 32);
    if (w < 0 || w >= p.ncols) {
      atomicOr(p.err, DEV_ERR_INDEX);
    } else {
      start = ld_gather64_i64(p.ptrs + w);
      const int64_t d = ld_gather64_i64(p.ptrs + w + 1) - start;
      if (d < 0 || d > 0x7fffffffll) atomicOr(p.err, DEV_ERR_INDEX);
      else deg = (uint32_t)d;
    }
    if (KIND != TCHGEO_SAMPLER_UNIFORM_REPLACE && k == 0 && deg > 0) atomicOr(p.err, DEV_ERR_PANIC);
    if (KIND == TCHGEO_SAMPLER_UNIFORM && deg > k) {
      nblocks = (deg - k + 3u) >> 2;
      heavy = nblocks > (uint32_t)SV_LIGHT_MAX;
    }
  }
  s_start[tid] = start;
  s_deg[tid] = deg;
  const uint32_t light = heavy ? 0u : nblocks;
  uint32_t choff, Q;
  BlockScan(scan_tmp).ExclusiveSum(light, choff, Q);
  s_choff[tid] = choff;
  for (uint32_t e = tid; e < (uint32_t)nn * k; e += SV_THREADS) s_slot[e] = 0u;
  if (tid < nn) {
    for (uint32_t c = 0; c < light; ++c) s_chown[choff + c] = (uint8_t)tid;
    if (heavy) s_heavy[atomicAdd(&s_nheavy, 1u)] = (uint8_t)tid;
  }
  __syncthreads();

  if (KIND == TCHGEO_SAMPLER_UNIFORM) {
    const uint32_t tag = TAG_RESERVOIR | (p.rel << 8);
    for (uint32_t q = tid; q < Q; q += SV_THREADS) {
      const uint32_t n = s_chown[q];
      const uint32_t c = q - s_choff[n];
      sv_block(philox4x32_10(s_pos[n], c, s_batch[n], tag, p.key0, p.key1), k + 4u * c, s_deg[n], k, s_slot + n * k);
    }
    const uint32_t nheavy = s_nheavy;
    for (uint32_t h = (uint32_t)(tid >> 5); h < nheavy; h += SV_THREADS / 32) {
      const uint32_t n = s_heavy[h];
      const uint32_t dn = s_deg[n], nb = (dn - k + 3u) >> 2;
      for (uint32_t c = lane; c < nb; c += 32)
        sv_block(philox4x32_10(s_pos[n], c, s_batch[n], tag, p.key0, p.key1), k + 4u * c, dn, k, s_slot + n * k);
    }
    __syncthreads();
  } else if (KIND == TCHGEO_SAMPLER_WEIGHTED) {
    for (int n = tid >> 5; n < nn; n += SV_THREADS / 32) {
      const uint32_t dn = s_deg[n];
      if (dn <= k) continue;
      const double* wp = p.weights + s_start[n];
      double carry = 0.0;
      for (uint32_t base = 0; base < dn; base += 32) {
        const uint32_t item = base + lane;
        const double w = item < dn ? __ldg(wp + item) : 0.0;
        double incl = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const double w_sum = carry + incl;
        if (item >= k && item < dn) {
          if (!(w_sum > 0.0)) {
            atomicOr(p.err, DEV_ERR_PANIC);
          } else {
            const Philox4 r = philox4x32_10(s_pos[n], item, s_batch[n], TAG_WEIGHTED | (p.rel << 8), p.key0, p.key1);
            const uint64_t u53 = ((uint64_t)r.x << 21) | (uint64_t)(r.y >> 11);
            const double u = __dmul_rn((double)u53, 1.0 / 9007199254740992.0);
            if (__dmul_rn(u, w_sum) < w) atomicMax(s_slot + n * k + __umulhi(r.z, k), item);
          }
        }
        carry = __shfl_sync(0xffffffffu, w_sum, 31);
      }
    }
    __syncthreads();
  }

  // answers: fanout slots per request, contiguous over the tile
  for (uint32_t e = tid; e < (uint32_t)nn * k; e += SV_THREADS) {
    const uint32_t n = e / k, s = e - n * k;
    const uint32_t dn = s_deg[n];
    const uint32_t cnt = KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE ? (dn > 0 ? k : 0u) : min(dn, k);
    int64_t id = -1, gp = -1;
    if (s < cnt) {
      uint32_t rel_ptr;
      if (KIND == TCHGEO_SAMPLER_UNIFORM_REPLACE) {
        const Philox4 r = philox4x32_10(s_pos[n], s >> 2, s_batch[n], TAG_REPLACE | (p.rel << 8), p.key0, p.key1);
        rel_ptr = __umulhi(pick4(r, s & 3u), dn);
      } else {
        const uint32_t st = s_slot[e];
        rel_ptr = st ? st : s;
      }
      const int64_t lp = s_start[n] + rel_ptr;
      id = ld_gather64_i64(p.indices + lp);
      gp = p.edge_base + lp;
    }
    if (p.out32 || p.peer_world > 0) {
      // compact rows are staged in shared memory and leave in row-major order below: consecutive lanes then write
      // consecutive words of consecutive rows, i.e. whole lines (what NVLink peer stores need: two interleaved
      // 4k-byte halves per row fill every packet only half)
      s_ans[n * 2u * k + s] = (int32_t)id;
      s_ans[n * 2u * k + k + s] = gp < 0 ? -1 : (int32_t)(gp - p.edge_base);
    } else {
      st_cs_i64(p.out_ids + (r0 + n) * p.out_stride + s, id);
      st_cs_i64(p.out_ptrs + (r0 + n) * p.out_stride + s, gp);
    }
  }
  if (p.out32 || p.peer_world > 0) {
    __syncthreads();
    const uint32_t w2 = 2u * k;
    for (uint32_t i = tid; i < (uint32_t)nn * w2; i += SV_THREADS) {
      const uint32_t n = i / w2;
      s_row[n][i - n * w2] = s_ans[i];
    }
  }
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

static tchgeo_status serve_launch(const int64_t* ptrs_local, const int64_t* indices_local, const double* weights_local,
                                  int64_t col_begin, int64_t ncols_local, int64_t edge_base, const int64_t* req_ids,
                                  const int64_t* req_meta, int64_t req_stride, int64_t n, int64_t fanout,
                                  int32_t sampler_kind, uint64_t seed, uint32_t rel, int64_t* out_ids, int64_t* out_ptrs,
                                  int64_t out_stride, int32_t* out32, uint32_t* err, cudaStream_t stream,
                                  int32_t peer_world = 0, void* const* peer_base = nullptr,
                                  const int64_t* peer_counts = nullptr, const int64_t* peer_row0 = nullptr) {
  TCHGEO_REQUIRE(n >= 0 && fanout >= 0 && fanout <= SV_MAX_TILE_SLOTS && ncols_local >= 0, "bad serve argument");
  TCHGEO_REQUIRE(sampler_kind >= 0 && sampler_kind <= 2 && err != nullptr, "bad serve argument");
  TCHGEO_REQUIRE(sampler_kind != TCHGEO_SAMPLER_WEIGHTED || weights_local != nullptr, "weighted serve without weights");
  // fanout 0: sampling with replacement yields nothing; the other samplers panic in the reference as soon as a
  // requested column is not empty (gen_range(0..0), src/utils/sampling.rs:19), which the kernel reports
  if (n == 0 || (fanout == 0 && sampler_kind == TCHGEO_SAMPLER_UNIFORM_REPLACE)) return TCHGEO_OK;
  TCHGEO_REQUIRE(ptrs_local && req_ids && req_meta && (fanout == 0 || out32 || peer_world > 0 || (out_ids && out_ptrs)),
                 "NULL pointer");
  ServeParams sp;
  sp.peer_world = peer_world;
  if (peer_world > 0) {
    TCHGEO_REQUIRE(peer_world <= SV_MAX_WORLD && peer_base && peer_counts && peer_row0, "bad peer table");
    int64_t at = 0;
    for (int q = 0; q < peer_world; ++q) {
      TCHGEO_REQUIRE(peer_counts[q] >= 0 && peer_row0[q] >= 0 && (peer_counts[q] == 0 || peer_base[q]), "bad peer table entry %d", q);
      sp.peer_base[q] = (int32_t*)peer_base[q];
      sp.peer_seg[q] = at;
      sp.peer_row0[q] = peer_row0[q];
      at += peer_counts[q];
    }
    sp.peer_seg[peer_world] = at;
    TCHGEO_REQUIRE(at == n, "peer segment sizes must add up to the number of requests");
  }
  sp.ptrs = ptrs_local; sp.indices = indices_local; sp.weights = weights_local;
  sp.req_ids = req_ids; sp.req_meta = req_meta; sp.req_stride = req_stride;
  sp.out_ids = out_ids; sp.out_ptrs = out_ptrs; sp.out_stride = out_stride; sp.out32 = out32;
  sp.err = err;
  sp.col_begin = col_begin; sp.ncols = ncols_local; sp.edge_base = edge_base; sp.n = n;
  sp.fanout = (int32_t)fanout;
  sp.tile_reqs = (int32_t)std::min<int64_t>(SV_THREADS, std::max<int64_t>(1, SV_MAX_TILE_SLOTS / std::max<int64_t>(fanout, 1)));
  sp.key0 = (uint32_t)seed; sp.key1 = (uint32_t)(seed >> 32); sp.rel = rel;
  const int64_t grid = (n + sp.tile_reqs - 1) / sp.tile_reqs;
  TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many requests for one launch");
  const bool compact = out32 != nullptr || peer_world > 0;
  const size_t smem = (size_t)sp.tile_reqs * fanout * (compact ? 12 : 4) + 16;  // slots (+ staged compact answer rows)
  {
    static bool configured[64] = {};  // per device; benign race: the attribute is idempotent
    int dev = 0;
    TCHGEO_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(serve_kernel<TCHGEO_SAMPLER_UNIFORM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(serve_kernel<TCHGEO_SAMPLER_UNIFORM_REPLACE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(serve_kernel<TCHGEO_SAMPLER_WEIGHTED>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  switch (sampler_kind) {
    case TCHGEO_SAMPLER_UNIFORM: serve_kernel<TCHGEO_SAMPLER_UNIFORM><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
    case TCHGEO_SAMPLER_UNIFORM_REPLACE: serve_kernel<TCHGEO_SAMPLER_UNIFORM_REPLACE><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
    default: serve_kernel<TCHGEO_SAMPLER_WEIGHTED><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_serve_requests(const int64_t* ptrs_local, const int64_t* indices_local,
                                               const double* weights_local, int64_t col_begin, int64_t ncols_local,
                                               int64_t edge_base, const int64_t* req_ids, const int64_t* req_meta,
                                               int64_t n, int64_t fanout, int32_t sampler_kind, uint64_t seed,
                                               uint32_t rel, int64_t* out_ids, int64_t* out_ptrs,
                                               int32_t* err_scratch, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(err_scratch != nullptr, "bad serve argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(err_scratch, 0, 4, stream));
  const tchgeo_status st = serve_launch(ptrs_local, indices_local, weights_local, col_begin, ncols_local, edge_base,
                                        req_ids, req_meta, 1, n, fanout, sampler_kind, seed, rel, out_ids, out_ptrs,
                                        fanout, nullptr, (uint32_t*)err_scratch, stream);
  if (st != TCHGEO_OK) return st;
  uint32_t herr = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, err_scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(herr);
}

// ---------------------------------------------------------------------------------------------
// Requester side of the partitioned path as a device pipeline (no host round trip inside a hop except the
// exchange of the per-owner request counts that sizes the all-to-all):
//   tchgeo_part_begin_hop : count the frontier per owning rank, then scatter (id, meta) request rows into the
//                           send buffer grouped by owner (order inside a group is arbitrary: answers come back in
//                           request order and are placed by the (batch, position) they carry)
//   tchgeo_serve_requests_rows : owner side on interleaved rows, asynchronous
//   tchgeo_part_finish_hop: per-node answer counts -> exclusive scan in frontier order -> the tree layout of
//                           neighbor_sampling.rs:210-218 in the caller's [B, capacity] buffers, new lengths
// ---------------------------------------------------------------------------------------------
namespace tchgeo {
namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_MAX_WORLD = 64;

struct PartFrontier {
  const int64_t* samples;     // [B, samples_stride]
  int64_t samples_stride;
  const int64_t* fr_begin;    // [B] or NULL (= 0)
  const int64_t* fr_end;      // [B]
  int64_t B, capF;            // capF: upper bound of the per-batch frontier size (launch geometry)
  int64_t cols_per_rank;
  int32_t world;
  uint32_t batch_base;
};

// frontier slot j of batch b (the kernels run on a 2-D grid: blockIdx.y = batch, x covers the frontier capacity)
__device__ __forceinline__ bool pt_node(const PartFrontier& f, int64_t b, int64_t j, int64_t& pos, int64_t& id) {
  if (b >= f.B || j >= f.capF) return false;
  const int64_t fb = f.fr_begin ? f.fr_begin[b] : 0;
  int64_t fe = f.fr_end[b];
  if (fe > f.samples_stride) fe = f.samples_stride;
  if (j >= fe - fb) return false;
  pos = fb + j;
  id = f.samples[b * f.samples_stride + pos];
  return true;
}

// owner(w) = min(w / cols_per_rank, world - 1); out-of-range ids go to the edge ranks, whose serve kernel reports them
__device__ __forceinline__ int pt_owner_of(int64_t id, int64_t cols_per_rank, int world) {
  int64_t o;
  if (id < 0) o = 0;
  else if (((uint64_t)id | (uint64_t)cols_per_rank) >> 32) o = id / cols_per_rank;
  else o = (int64_t)((uint32_t)id / (uint32_t)cols_per_rank);  // the common case: a 32-bit division
  if (o >= world) o = world - 1;
  return (int)o;
}
__device__ __forceinline__ int pt_owner(const PartFrontier& f, int64_t id) {
  return pt_owner_of(id, f.cols_per_rank, f.world);
}

// warp-aggregated histogram update: lanes with the same owner elect a leader that adds the group's size once;
// returns this lane's rank inside its owner's bucket of the CTA (frontiers of few owners would otherwise
// serialise 256 shared-memory atomics on one or two addresses)
__device__ __forceinline__ unsigned int pt_hist_add(unsigned int* hist, bool ok, int owner) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned m = __match_any_sync(0xffffffffu, ok ? owner : -1);
  const int leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (ok && (int)lane == leader) base = atomicAdd(&hist[owner], (unsigned int)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (unsigned int)__popc(m & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(PT_THREADS) part_count_kernel(const PartFrontier f, unsigned long long* counts,
                                                               uint32_t* err) {
  __shared__ unsigned int hist[PT_MAX_WORLD];
  if (threadIdx.x < PT_MAX_WORLD) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t b = blockIdx.y, j = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
  int64_t pos, id = 0;
  if (j == 0 && f.fr_end[b] - (f.fr_begin ? f.fr_begin[b] : 0) > f.capF)  // the frontier must fit the launch geometry
    atomicOr(err, DEV_ERR_CAPACITY);
  const bool ok = pt_node(f, b, j, pos, id);
  pt_hist_add(hist, ok, ok ? pt_owner(f, id) : 0);
  __syncthreads();
  if (threadIdx.x < f.world && hist[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
}

// The request exchange as peer-memory stores: the send buffer is grouped by owner, so group o is one contiguous run of
// 16-byte rows that goes to rows [row0[o], row0[o] + counts[o]) of owner o's request buffer -- the rows the request
// all-to-all would have delivered it to.  Consecutive threads move consecutive rows: whole lines per warp (scattering
// the rows to the peers one by one from part_scatter_kernel filled every NVLink packet with 16 bytes and took 6 ms of
// an 8-GPU step).
struct PeerReq {
  int32_t world;                // 0: off
  int64_t* base[PT_MAX_WORLD];  // owner o's request buffer [rows, 2]
  int64_t row0[PT_MAX_WORLD];   // first row of this rank's segment in owner o's buffer
};

__global__ void __launch_bounds__(PT_THREADS) part_scatter_kernel(const PartFrontier f, const unsigned long long* counts,
                                                                 unsigned long long* cursor, int64_t* req,
                                                                 int32_t* slot_of) {
  __shared__ unsigned int hist[PT_MAX_WORLD];
  __shared__ unsigned long long base[PT_MAX_WORLD];    // row in the local send buffer
  if (threadIdx.x < PT_MAX_WORLD) hist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t b = blockIdx.y, j = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
  int64_t pos = 0, id = 0;
  const bool ok = pt_node(f, b, j, pos, id);
  const int o = ok ? pt_owner(f, id) : 0;
  const unsigned int rank = pt_hist_add(hist, ok, o);
  __syncthreads();
  if (threadIdx.x < f.world) {
    unsigned long long off = 0;  // exclusive offset of this owner's group in the send buffer
    for (int r = 0; r < threadIdx.x; ++r) off += counts[r];
    const unsigned int h = hist[threadIdx.x];
    base[threadIdx.x] = off + (h ? atomicAdd(cursor + threadIdx.x, (unsigned long long)h) : 0ull);
  }
  __syncthreads();
  if (ok) {
    const unsigned long long q = base[o] + rank;
    const uint64_t meta = ((uint64_t)(f.batch_base + (uint32_t)b) << 32) | (uint64_t)(uint32_t)pos;
    const longlong2 row = make_longlong2((long long)id, (long long)meta);
    *reinterpret_cast<longlong2*>(req + 2 * q) = row;  // grouped by owner; part_put_kernel ships the groups
    slot_of[b * f.capF + j] = (int32_t)q;  // where this frontier node's answer will be found (frontier order, coalesced)
  }
}

__global__ void __launch_bounds__(PT_THREADS) part_put_kernel(const int64_t* __restrict__ req,
                                                             const unsigned long long* __restrict__ counts,
                                                             const PeerReq peer) {
  __shared__ unsigned long long seg[PT_MAX_WORLD + 1];
  if (threadIdx.x == 0) {
    unsigned long long at = 0;
    for (int o = 0; o < peer.world; ++o) {
      seg[o] = at;
      at += counts[o];
    }
    seg[peer.world] = at;
  }
  __syncthreads();
  const unsigned long long total = seg[peer.world];
  for (unsigned long long q = (unsigned long long)blockIdx.x * PT_THREADS + threadIdx.x; q < total;
       q += (unsigned long long)gridDim.x * PT_THREADS) {
    int o = 0;
    while (o + 1 < peer.world && q >= seg[o + 1]) ++o;
    const longlong2 row = *reinterpret_cast<const longlong2*>(req + 2 * q);
    *reinterpret_cast<longlong2*>(peer.base[o] + 2 * (peer.row0[o] + (int64_t)(q - seg[o]))) = row;
  }
}

struct PartFinish {
  const int64_t* req;      // [F, 2] the requests this rank sent (same order as the answers)
  const int32_t* ans;      // [F, 2k] int32 rows: k ids then k LOCAL csc positions of the owner, -1 padded
  const int64_t* edge_base;  // [world] CSC entries owned by lower ranks: global position = edge_base[owner] + local
  int64_t cols_per_rank;
  int32_t world;
  int64_t F;
  int32_t k;
  const int64_t* fr_begin; // [B] or NULL
  const int64_t* fr_end;   // [B]
  int64_t B, capF;
  const int32_t* slot_of;  // [B * capF] request slot of every frontier node (written by tchgeo_part_begin_hop)
  int32_t* fcnt;           // [B * capF + 1] answers per frontier node (padding = 0), then its exclusive scan in place
  const int64_t* node_len_in;  // [B]
  const int64_t* edge_len_in;  // [B]
  int64_t* node_len_out;
  int64_t* edge_len_out;
  int64_t* samples; int64_t samples_stride;
  int64_t* rows; int64_t* cols; int64_t* eidx; int64_t edges_stride;
  uint32_t* err;
};

// Both kernels walk the frontier in ORDER (2-D grid: blockIdx.y = batch) and fetch the node's answer row through
// slot_of: the output arrays are then written front to back in full lines whatever the number of owners (walking the
// answers in arrival order instead sweeps the outputs once per owner and leaves every line half written).
__global__ void __launch_bounds__(PT_THREADS) part_cnt_kernel(const PartFinish p) {
  const int64_t b = blockIdx.y, j = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
  if (j >= p.capF) return;
  const int64_t fb = p.fr_begin ? p.fr_begin[b] : 0;
  int c = 0;
  if (j < p.fr_end[b] - fb) {
    const int32_t* a = p.ans + (int64_t)p.slot_of[b * p.capF + j] * 2 * p.k + p.k;
    for (int s = 0; s < p.k; ++s) c += a[s] >= 0 ? 1 : 0;  // valid slots form a prefix
  }
  p.fcnt[b * p.capF + j] = c;
  if (b == p.B - 1 && j == p.capF - 1) p.fcnt[p.B * p.capF] = 0;
}

__global__ void __launch_bounds__(PT_THREADS) part_len_kernel(const PartFinish p) {
  const int64_t b = (int64_t)blockIdx.x * PT_THREADS + threadIdx.x;
  if (b >= p.B) return;
  const int64_t total = (int64_t)p.fcnt[(b + 1) * p.capF] - (int64_t)p.fcnt[b * p.capF];
  const int64_t n = p.node_len_in[b] + total, e = p.edge_len_in[b] + total;
  p.node_len_out[b] = n;
  p.edge_len_out[b] = e;
  if (n > p.samples_stride || e > p.edges_stride) atomicOr(p.err, DEV_ERR_CAPACITY);
}

__global__ void __launch_bounds__(PT_THREADS) part_emit_kernel(const PartFinish p) {
  const int64_t b = blockIdx.y;
  const uint32_t t = blockIdx.x * PT_THREADS + threadIdx.x;   // < capF * k < 2^31 (checked by the host)
  const uint32_t j = t / (uint32_t)p.k;
  const int s = (int)(t - j * (uint32_t)p.k);
  if ((int64_t)j >= p.capF) return;
  const int64_t fb = p.fr_begin ? p.fr_begin[b] : 0;
  if ((int64_t)j >= p.fr_end[b] - fb) return;
  const int64_t idx = b * p.capF + j;
  const int32_t first = p.fcnt[idx];
  if (s >= p.fcnt[idx + 1] - first) return;   // only this node's valid slots
  const int64_t q = p.slot_of[idx];
  const int32_t* a = p.ans + q * 2 * p.k;
  const int64_t off = (int64_t)first - (int64_t)p.fcnt[b * p.capF] + s;
  const int64_t ni = p.node_len_in[b] + off, ei = p.edge_len_in[b] + off;
  if (ni >= p.samples_stride || ei >= p.edges_stride) return;  // flagged by part_len_kernel
  const int owner = pt_owner_of(p.req[2 * q], p.cols_per_rank, p.world);
  st_cs_i64(p.samples + b * p.samples_stride + ni, (int64_t)a[s]);
  st_cs_i64(p.rows + b * p.edges_stride + ei, ni);        // index of the appended node, neighbor_sampling.rs:213-216
  st_cs_i64(p.cols + b * p.edges_stride + ei, fb + j);    // index of the frontier node
  st_cs_i64(p.eidx + b * p.edges_stride + ei, p.edge_base[owner] + a[p.k + s]);  // global CSC position
}

struct PartWs {  // layout of the per-plan hop workspace shared by begin_hop (slot_of) and finish_hop
  size_t off_fcnt, off_slot, off_cub, cub_bytes, total;
};
bool part_ws_layout(int64_t num_batches, int64_t frontier_cap, PartWs& w) {
  const int64_t n = num_batches * frontier_cap + 1;
  if (n <= 0 || n >= ((int64_t)1 << 31)) return false;
  size_t cub_bytes = 0;
  if (cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, n) != cudaSuccess)
    return false;
  const size_t arr = ((size_t)n * 4 + 255) / 256 * 256;
  w.off_fcnt = 256;
  w.off_slot = w.off_fcnt + arr;
  w.off_cub = w.off_slot + arr;
  w.cub_bytes = cub_bytes;
  w.total = w.off_cub + (cub_bytes + 255) / 256 * 256;
  return true;
}

}  // namespace
}  // namespace tchgeo

static tchgeo_status part_begin(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap, int64_t cols_per_rank,
                                int32_t world, uint32_t batch_base, int64_t* counts, int64_t* cursor, int64_t* req,
                                int32_t* err_word, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                bool do_count, bool do_scatter, void* const* peer_req, const int64_t* peer_row0) {
  TCHGEO_REQUIRE(world >= 1 && world <= PT_MAX_WORLD, "world size must be in [1, 64]");
  TCHGEO_REQUIRE(err_word != nullptr, "NULL pointer");
  TCHGEO_REQUIRE(num_batches >= 0 && frontier_cap >= 0 && cols_per_rank >= 1 && samples_stride >= 0, "bad argument");
  TCHGEO_REQUIRE(counts && cursor, "NULL pointer");
  if (do_count) {
    TCHGEO_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)world * 8, stream));
    TCHGEO_CUDA_CHECK(cudaMemsetAsync(cursor, 0, (size_t)world * 8, stream));
  }
  if (num_batches == 0 || frontier_cap == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(samples && fr_end && (req || !do_scatter), "NULL pointer");
  TCHGEO_REQUIRE(num_batches <= 65535, "at most 65535 batches per call");
  PartWs W;
  TCHGEO_REQUIRE(part_ws_layout(num_batches, frontier_cap, W), "frontier too large for one call");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= W.total, "workspace too small: need %zu bytes", W.total);
  int32_t* slot_of = (int32_t*)((char*)workspace + W.off_slot);
  const int64_t gx = (frontier_cap + PT_THREADS - 1) / PT_THREADS;
  TCHGEO_REQUIRE(gx < ((int64_t)1 << 31), "frontier too large for one launch");
  const dim3 grid((unsigned)gx, (unsigned)num_batches);
  PartFrontier f;
  f.samples = samples; f.samples_stride = samples_stride; f.fr_begin = fr_begin; f.fr_end = fr_end;
  f.B = num_batches; f.capF = frontier_cap; f.cols_per_rank = cols_per_rank; f.world = world; f.batch_base = batch_base;
  if (do_count) {
    part_count_kernel<<<grid, PT_THREADS, 0, stream>>>(f, (unsigned long long*)counts, (uint32_t*)err_word);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  if (do_scatter) {
    PeerReq peer;
    peer.world = 0;
    if (peer_req) {
      TCHGEO_REQUIRE(peer_row0 != nullptr, "peer_row0 is NULL");
      peer.world = world;
      for (int o = 0; o < world; ++o) {
        TCHGEO_REQUIRE(peer_req[o] != nullptr && peer_row0[o] >= 0, "bad peer request table entry %d", o);
        peer.base[o] = (int64_t*)peer_req[o];
        peer.row0[o] = peer_row0[o];
      }
    }
    part_scatter_kernel<<<grid, PT_THREADS, 0, stream>>>(f, (const unsigned long long*)counts,
                                                        (unsigned long long*)cursor, req, slot_of);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    if (peer.world > 0) {
      const int64_t rows_max = num_batches * frontier_cap;
      const unsigned pgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((rows_max + PT_THREADS - 1) / PT_THREADS, 148 * 8));
      part_put_kernel<<<pgrid, PT_THREADS, 0, stream>>>(req, (const unsigned long long*)counts, peer);
      TCHGEO_CUDA_CHECK(cudaGetLastError());
    }
  }
  return TCHGEO_OK;
}

extern "C" tchgeo_status tchgeo_part_begin_hop(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                               const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                               int64_t cols_per_rank, int32_t world, uint32_t batch_base,
                                               int64_t* counts, int64_t* cursor, int64_t* req, int32_t* err_word,
                                               void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  return part_begin(samples, samples_stride, fr_begin, fr_end, num_batches, frontier_cap, cols_per_rank, world, batch_base,
                    counts, cursor, req, err_word, workspace, workspace_bytes, (cudaStream_t)stream_, true, true, nullptr,
                    nullptr);
}

extern "C" tchgeo_status tchgeo_part_count_hop(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                               const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                               int64_t cols_per_rank, int32_t world, int64_t* counts, int64_t* cursor,
                                               int32_t* err_word, void* workspace, size_t workspace_bytes,
                                               tchgeo_stream stream_) {
  return part_begin(samples, samples_stride, fr_begin, fr_end, num_batches, frontier_cap, cols_per_rank, world, 0, counts,
                    cursor, nullptr, err_word, workspace, workspace_bytes, (cudaStream_t)stream_, true, false, nullptr,
                    nullptr);
}

extern "C" tchgeo_status tchgeo_part_scatter_hop(const int64_t* samples, int64_t samples_stride, const int64_t* fr_begin,
                                                 const int64_t* fr_end, int64_t num_batches, int64_t frontier_cap,
                                                 int64_t cols_per_rank, int32_t world, uint32_t batch_base,
                                                 const int64_t* counts, int64_t* cursor, int64_t* req,
                                                 void* const* peer_req, const int64_t* peer_row0, int32_t* err_word,
                                                 void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  return part_begin(samples, samples_stride, fr_begin, fr_end, num_batches, frontier_cap, cols_per_rank, world, batch_base,
                    const_cast<int64_t*>(counts), cursor, req, err_word, workspace, workspace_bytes, (cudaStream_t)stream_,
                    false, true, peer_req, peer_row0);
}

extern "C" tchgeo_status tchgeo_serve_requests_rows(const int64_t* ptrs_local, const int64_t* indices_local,
                                                    const double* weights_local, int64_t col_begin, int64_t ncols_local,
                                                    int64_t nnz_local, const int64_t* req, int64_t n, int64_t fanout,
                                                    int32_t sampler_kind, uint64_t seed, uint32_t rel, int32_t* ans,
                                                    int32_t* err_word, tchgeo_stream stream_) {
  // compact answers: neighbour ids and LOCAL csc positions travel as int32 (half the all-to-all bytes)
  TCHGEO_REQUIRE(nnz_local >= 0 && nnz_local < ((int64_t)1 << 31), "a rank's share of the CSC must stay below 2^31 entries");
  return serve_launch(ptrs_local, indices_local, weights_local, col_begin, ncols_local, 0, req, req ? req + 1 : req, 2, n,
                      fanout, sampler_kind, seed, rel, nullptr, nullptr, 0, ans, (uint32_t*)err_word,
                      (cudaStream_t)stream_);
}

extern "C" tchgeo_status tchgeo_serve_requests_rows_peer(const int64_t* ptrs_local, const int64_t* indices_local,
                                                         const double* weights_local, int64_t col_begin,
                                                         int64_t ncols_local, int64_t nnz_local, const int64_t* req,
                                                         int64_t n, int64_t fanout, int32_t sampler_kind, uint64_t seed,
                                                         uint32_t rel, int32_t world, void* const* peer_ans,
                                                         const int64_t* recv_counts, const int64_t* peer_row0,
                                                         int32_t* err_word, tchgeo_stream stream_) {
  // The owner's half of the answer exchange fused into the serve kernel: every answer row is stored straight into the
  // REQUESTER's answer buffer over NVLink peer memory, at the row the all-to-all would have delivered it to.
  TCHGEO_REQUIRE(nnz_local >= 0 && nnz_local < ((int64_t)1 << 31), "a rank's share of the CSC must stay below 2^31 entries");
  TCHGEO_REQUIRE(world >= 1, "world must be positive");
  return serve_launch(ptrs_local, indices_local, weights_local, col_begin, ncols_local, 0, req, req ? req + 1 : req, 2, n,
                      fanout, sampler_kind, seed, rel, nullptr, nullptr, 0, nullptr, (uint32_t*)err_word,
                      (cudaStream_t)stream_, world, peer_ans, recv_counts, peer_row0);
}

extern "C" size_t tchgeo_part_hop_workspace_bytes(int64_t num_batches, int64_t frontier_cap) {
  PartWs W;
  return part_ws_layout(num_batches, frontier_cap, W) ? W.total : 0;
}

extern "C" tchgeo_status tchgeo_part_finish_hop(const int64_t* req, const int32_t* ans, int64_t num_requests,
                                                int64_t fanout, const int64_t* owner_edge_base, int64_t cols_per_rank,
                                                int32_t world, const int64_t* fr_begin, const int64_t* fr_end,
                                                int64_t num_batches, int64_t frontier_cap, const int64_t* node_len_in,
                                                const int64_t* edge_len_in, int64_t* node_len_out, int64_t* edge_len_out,
                                                int64_t* samples, int64_t samples_stride, int64_t* rows, int64_t* cols,
                                                int64_t* edge_index, int64_t edges_stride, int32_t* err_word,
                                                void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_requests >= 0 && fanout >= 0 && fanout < (1 << 20) && num_batches >= 0 && frontier_cap >= 0,
                 "bad argument");
  TCHGEO_REQUIRE(num_batches <= 65535, "at most 65535 batches per call");
  TCHGEO_REQUIRE(node_len_in && edge_len_in && node_len_out && edge_len_out && err_word && owner_edge_base && fr_end,
                 "NULL pointer");
  TCHGEO_REQUIRE(world >= 1 && world <= PT_MAX_WORLD && cols_per_rank >= 1, "bad partition");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_batches == 0) return TCHGEO_OK;
  if (frontier_cap == 0 || fanout == 0) {  // nothing can be appended: the lengths carry over
    TCHGEO_CUDA_CHECK(cudaMemcpyAsync(node_len_out, node_len_in, (size_t)num_batches * 8, cudaMemcpyDeviceToDevice, stream));
    TCHGEO_CUDA_CHECK(cudaMemcpyAsync(edge_len_out, edge_len_in, (size_t)num_batches * 8, cudaMemcpyDeviceToDevice, stream));
    return TCHGEO_OK;
  }
  PartWs W;
  TCHGEO_REQUIRE(part_ws_layout(num_batches, frontier_cap, W), "frontier too large for one call");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= W.total, "workspace too small: need %zu bytes", W.total);
  TCHGEO_REQUIRE(frontier_cap * fanout < ((int64_t)1 << 31), "frontier_cap * fanout must stay below 2^31");
  // the per-node answer counts of ALL batches are scanned together in int32
  TCHGEO_REQUIRE(num_batches * frontier_cap * fanout < ((int64_t)1 << 31),
                 "num_batches * frontier_cap * fanout must stay below 2^31: use fewer batches per call");
  TCHGEO_REQUIRE(num_requests == 0 || (req && ans), "NULL pointer");
  TCHGEO_REQUIRE(samples && rows && cols && edge_index, "NULL pointer");
  const int64_t n = num_batches * frontier_cap + 1;
  PartFinish p;
  p.req = req; p.ans = ans; p.edge_base = owner_edge_base; p.cols_per_rank = cols_per_rank; p.world = world;
  p.F = num_requests; p.k = (int32_t)fanout; p.fr_begin = fr_begin; p.fr_end = fr_end;
  p.B = num_batches; p.capF = frontier_cap;
  p.slot_of = (const int32_t*)((char*)workspace + W.off_slot);
  p.fcnt = (int32_t*)((char*)workspace + W.off_fcnt);
  p.node_len_in = node_len_in; p.edge_len_in = edge_len_in; p.node_len_out = node_len_out; p.edge_len_out = edge_len_out;
  p.samples = samples; p.samples_stride = samples_stride; p.rows = rows; p.cols = cols; p.eidx = edge_index;
  p.edges_stride = edges_stride; p.err = (uint32_t*)err_word;
  const dim3 gcnt((unsigned)((frontier_cap + PT_THREADS - 1) / PT_THREADS), (unsigned)num_batches);
  part_cnt_kernel<<<gcnt, PT_THREADS, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  size_t cub_bytes = W.cub_bytes;
  TCHGEO_CUDA_CHECK(cub::DeviceScan::ExclusiveSum((char*)workspace + W.off_cub, cub_bytes, (const int32_t*)p.fcnt, p.fcnt, n,
                                                  stream));
  part_len_kernel<<<(unsigned)((num_batches + PT_THREADS - 1) / PT_THREADS), PT_THREADS, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  const dim3 gemit((unsigned)((frontier_cap * fanout + PT_THREADS - 1) / PT_THREADS), (unsigned)num_batches);
  part_emit_kernel<<<gemit, PT_THREADS, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}
