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

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int SV_THREADS = 128;
constexpr int SV_LIGHT_MAX = 32;
constexpr int SV_MAX_TILE_SLOTS = 8192;

struct ServeParams {
  const int64_t* ptrs;      // local colptr [ncols+1], rebased to the local indices array
  const int64_t* indices;   // local row_indices
  const double* weights;    // local weights or NULL
  const int64_t* req_ids;
  const int64_t* req_meta;  // (batch << 32) | pos
  int64_t* out_ids;
  int64_t* out_ptrs;
  uint32_t* err;
  int64_t col_begin, ncols, edge_base, n;
  int32_t fanout, tile_reqs;
  uint32_t key0, key1, rel;
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
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(dyn_smem);  // [tile_reqs * fanout]

  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t k = (uint32_t)p.fanout;
  const int64_t r0 = (int64_t)blockIdx.x * p.tile_reqs;
  const int nn = (int)min((int64_t)p.tile_reqs, p.n - r0);
  if (tid == 0) s_nheavy = 0u;

  uint32_t deg = 0, nblocks = 0;
  int64_t start = 0;
  bool heavy = false;
  if (tid < nn) {
    const int64_t w = p.req_ids[r0 + tid] - p.col_begin;
    const int64_t meta = p.req_meta[r0 + tid];
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
    st_cs_i64(p.out_ids + r0 * k + e, id);
    st_cs_i64(p.out_ptrs + r0 * k + e, gp);
  }
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_serve_requests(const int64_t* ptrs_local, const int64_t* indices_local,
                                               const double* weights_local, int64_t col_begin, int64_t ncols_local,
                                               int64_t edge_base, const int64_t* req_ids, const int64_t* req_meta,
                                               int64_t n, int64_t fanout, int32_t sampler_kind, uint64_t seed,
                                               uint32_t rel, int64_t* out_ids, int64_t* out_ptrs,
                                               int32_t* err_scratch, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n >= 0 && fanout >= 0 && fanout <= SV_MAX_TILE_SLOTS && ncols_local >= 0, "bad serve argument");
  TCHGEO_REQUIRE(sampler_kind >= 0 && sampler_kind <= 2 && err_scratch != nullptr, "bad serve argument");
  TCHGEO_REQUIRE(sampler_kind != TCHGEO_SAMPLER_WEIGHTED || weights_local != nullptr, "weighted serve without weights");
  if (n == 0 || fanout == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(ptrs_local && req_ids && req_meta && out_ids && out_ptrs, "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  ServeParams sp;
  sp.ptrs = ptrs_local; sp.indices = indices_local; sp.weights = weights_local;
  sp.req_ids = req_ids; sp.req_meta = req_meta; sp.out_ids = out_ids; sp.out_ptrs = out_ptrs;
  sp.err = (uint32_t*)err_scratch;
  sp.col_begin = col_begin; sp.ncols = ncols_local; sp.edge_base = edge_base; sp.n = n;
  sp.fanout = (int32_t)fanout;
  sp.tile_reqs = (int32_t)std::min<int64_t>(SV_THREADS, std::max<int64_t>(1, SV_MAX_TILE_SLOTS / fanout));
  sp.key0 = (uint32_t)seed; sp.key1 = (uint32_t)(seed >> 32); sp.rel = rel;
  const int64_t grid = (n + sp.tile_reqs - 1) / sp.tile_reqs;
  TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many requests for one launch");
  const size_t smem = (size_t)sp.tile_reqs * fanout * 4 + 16;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(err_scratch, 0, 4, stream));
  switch (sampler_kind) {
    case TCHGEO_SAMPLER_UNIFORM: serve_kernel<TCHGEO_SAMPLER_UNIFORM><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
    case TCHGEO_SAMPLER_UNIFORM_REPLACE: serve_kernel<TCHGEO_SAMPLER_UNIFORM_REPLACE><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
    default: serve_kernel<TCHGEO_SAMPLER_WEIGHTED><<<(unsigned)grid, SV_THREADS, smem, stream>>>(sp); break;
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t herr = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, err_scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(herr);
}
