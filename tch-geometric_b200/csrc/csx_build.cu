// edge_index (COO) -> CSC / CSR on sm_100a, and ind2ptr.
//
// Replaces src/data/storage.rs:67-101 (ind2ptr) and :103-127 (argsort(col*N + row), ind2ptr,
// gathers) of the reference.  The sort key packs (major << minor_bits) | minor, which orders
// exactly like major*size_minor + minor because minor < size_minor <= 2^minor_bits, and only the
// key's significant bits are radix-sorted (44 bits for the products-shaped graph instead of 64).
// The value carried through the sort is the 32-bit edge id (perm); `indices` and the major ids are
// decoded from the sorted keys, so no random row[perm] / col[perm] gathers are needed.
// The device-wide radix sort itself is CUB's DeviceRadixSort (library code, like cuBLAS for a plain
// GEMM); key build, decode and ind2ptr are the kernels below.
//
// Second form ("partition form", the default where it applies; TCHGEO_CSX_SORT=cub selects the radix form): no
// device-wide sort.  The edges are split ONCE into
// buckets of 2^low_bits consecutive major ids (a few thousand edges each: tile histograms, a scan over tiles, a
// scatter of packed 8-byte (key, edge id) pairs), and one CTA per bucket finishes its bucket in shared memory:
// counting sort by major id, then every column's few dozen minor ids are sorted by a warp (bitonic network in
// registers for <= 32 entries, in shared memory above).  Global traffic is three passes over the edges instead
// of the radix sort's fourteen.  It applies when the buckets fit (see pt_plan / PT_CAP) and falls back to the
// radix form otherwise.
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int CSX_THREADS = 256;
constexpr bool PT_DEFAULT_ON = true;    // partition form where it applies; TCHGEO_CSX_SORT=cub | partition overrides

__host__ __device__ inline int bits_for(int64_t n) {  // bits needed for values in [0, n)
  int b = 0;
  while (b < 63 && ((int64_t)1 << b) < n) ++b;
  return b < 1 ? 1 : b;
}

__global__ void __launch_bounds__(CSX_THREADS) build_keys_kernel(const int64_t* __restrict__ major,
                                                                const int64_t* __restrict__ minor, int64_t E,
                                                                int64_t n_major, int64_t n_minor, int minor_bits,
                                                                uint64_t* __restrict__ keys,
                                                                uint32_t* __restrict__ vals, uint32_t* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const int64_t a = major[e], b = minor[e];
    if (a < 0 || a >= n_major || b < 0 || b >= n_minor) atomicOr(err, DEV_ERR_INDEX);
    keys[e] = ((uint64_t)a << minor_bits) | (uint64_t)b;
    vals[e] = (uint32_t)e;
  }
}

// sorted keys -> indices (minor ids), perm (i64) and ptrs (ind2ptr of the major ids, storage.rs:67-101)
__global__ void __launch_bounds__(CSX_THREADS) decode_kernel(const uint64_t* __restrict__ keys,
                                                            const uint32_t* __restrict__ vals, int64_t E,
                                                            int64_t n_major, int minor_bits,
                                                            int64_t* __restrict__ ptrs, int64_t* __restrict__ indices,
                                                            int64_t* __restrict__ perm) {
  const uint64_t mask = (minor_bits >= 64) ? ~0ull : ((1ull << minor_bits) - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const uint64_t key = keys[e];
    const int64_t maj = (int64_t)(key >> minor_bits);
    indices[e] = (int64_t)(key & mask);
    perm[e] = (int64_t)vals[e];
    // out[i] = e for every i in (major[e-1], major[e]]; out[0..=major[0]] = 0
    int64_t lo = (e == 0) ? 0 : (int64_t)(keys[e - 1] >> minor_bits) + 1;
    if (maj >= n_major) continue;  // flagged by build_keys_kernel
    for (int64_t i = lo; i <= maj; ++i) ptrs[i] = e;
    if (e == E - 1)
      for (int64_t i = maj + 1; i <= n_major; ++i) ptrs[i] = E;
  }
}

__global__ void __launch_bounds__(CSX_THREADS) ind2ptr_kernel(const int64_t* __restrict__ ind, int64_t numel,
                                                             int64_t m, int64_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (numel == 0) {  // storage.rs:78-80
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= m; i += stride) out[i] = 0;
    return;
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < numel; e += stride) {
    const int64_t cur = ind[e];
    int64_t lo = (e == 0) ? 0 : ind[e - 1] + 1;
    if (lo < 0) lo = 0;
    const int64_t hi = cur < m ? cur : m;
    for (int64_t i = lo; i <= hi; ++i) out[i] = e;
    if (e == numel - 1)
      for (int64_t i = (cur < -1 ? -1 : cur) + 1; i <= m; ++i) out[i] = numel;
  }
}

__global__ void fill_zero_i64_kernel(int64_t* p, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0;
}

struct CsxLayout {
  size_t off_err, off_keys0, off_keys1, off_vals0, off_vals1, off_cub, cub_bytes, total;
};

inline size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

cudaError_t csx_layout(int64_t E, int end_bit, CsxLayout& L) {
  size_t cub_bytes = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, (int64_t)(E > 0 ? E : 1), 0, end_bit);
  if (e != cudaSuccess) return e;
  const size_t n = (size_t)(E > 0 ? E : 1);
  L.off_err = 0;
  L.off_keys0 = 256;
  L.off_keys1 = L.off_keys0 + align_up(n * 8);
  L.off_vals0 = L.off_keys1 + align_up(n * 8);
  L.off_vals1 = L.off_vals0 + align_up(n * 4);
  L.off_cub = L.off_vals1 + align_up(n * 4);
  L.cub_bytes = cub_bytes;
  L.total = L.off_cub + align_up(cub_bytes) + 256;
  return cudaSuccess;
}

inline unsigned grid_for(int64_t n) {
  int64_t g = (n + CSX_THREADS - 1) / CSX_THREADS;
  const int64_t cap = 148 * 16;  // grid-stride loops: 16 resident 256-thread CTAs' worth per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}


// =================================================================================================
// Partition form
// =================================================================================================
constexpr int PT_THREADS = 1024;                 // count / scatter kernels
constexpr int PT_ITEMS = 32;
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;   // edges per CTA of the count / scatter kernels
constexpr int PT_GROUP = 8;                      // edges requested before the first is used
constexpr int PT_NB_MAX = 16384;                 // buckets (shared histogram of the count / scatter kernels: 64 KB)
constexpr int PT_CAP = 12288;                    // edges one bucket may hold (96 KB of 8-byte pairs, two CTAs per SM)
constexpr int PT_MAX_LOW_BITS = 10;              // <= 1024 major ids per bucket
constexpr int PT_BTHREADS = 512;                 // bucket kernel
constexpr int PT_SEGS = 64;                      // the scan over tiles runs per segment of tiles
constexpr int PT_WARP_MAX = 1024;                // longer columns are sorted by the whole CTA
constexpr int PT_HEAVY_MAX = 16;                 // > PT_CAP / PT_WARP_MAX

struct PtParams {
  const int64_t* major;
  const int64_t* minor;
  int64_t E, n_major, n_minor;
  int low_bits, minor_bits, nb, tiles, tiles_per_seg, segs;
  uint32_t* tile_hist;  // [tiles][nb] edges of the tile per bucket, then (seg prefix) those of earlier tiles of the segment
  uint32_t* seg_tot;    // [segs][nb] edges of the segment per bucket, then those of earlier segments
  uint32_t* tot;        // [nb] edges per bucket
  uint32_t* base;       // [nb] first slot of the bucket
  uint32_t* stats;      // [0] DEV_ERR_* bits, [1] largest bucket
  uint2* pairs;         // [E] (major low bits << minor_bits | minor, edge id), grouped by bucket
  int64_t* ptrs;
  int64_t* indices;
  int64_t* perm;
};

// ---- tile histograms ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_THREADS) pt_count_kernel(const PtParams p) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_hist = s_dyn;
  const int tid = threadIdx.x;
  for (int j = tid; j < p.nb; j += PT_THREADS) s_hist[j] = 0u;
  __syncthreads();
  const int64_t i0 = (int64_t)blockIdx.x * PT_TILE;
  bool bad = false;
#pragma unroll 1
  for (int u0 = 0; u0 < PT_ITEMS; u0 += PT_GROUP) {
    int64_t a[PT_GROUP];
#pragma unroll
    for (int u = 0; u < PT_GROUP; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * PT_THREADS + tid;
      a[u] = i < p.E ? __ldg(p.major + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < PT_GROUP; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * PT_THREADS + tid;
      if (i >= p.E) continue;
      if ((uint64_t)a[u] >= (uint64_t)p.n_major) bad = true;   // no slot for it: the scatter kernel skips it as well
      else atomicAdd(&s_hist[(uint32_t)(a[u] >> p.low_bits)], 1u);
    }
  }
  if (bad) atomicOr(p.stats, DEV_ERR_INDEX);
  __syncthreads();
  uint32_t* row = p.tile_hist + (size_t)blockIdx.x * p.nb;
  for (int j = tid; j < p.nb; j += PT_THREADS) row[j] = s_hist[j];
}

// ---- scan over tiles, per bucket: inside every segment of tiles, then over the segments, then over the buckets -------
__global__ void __launch_bounds__(256) pt_seg_prefix_kernel(const PtParams p) {
  const int j = blockIdx.x * 256 + threadIdx.x, seg = blockIdx.y;
  if (j >= p.nb) return;
  const int t0 = seg * p.tiles_per_seg;
  const int t1 = min(t0 + p.tiles_per_seg, p.tiles);
  uint32_t run = 0u;
  int t = t0;
  for (; t + 4 <= t1; t += 4) {   // four independent loads in flight
    uint32_t* c = p.tile_hist + (size_t)t * p.nb + j;
    const uint32_t v0 = c[0], v1 = c[(size_t)p.nb], v2 = c[2 * (size_t)p.nb], v3 = c[3 * (size_t)p.nb];
    c[0] = run; run += v0;
    c[(size_t)p.nb] = run; run += v1;
    c[2 * (size_t)p.nb] = run; run += v2;
    c[3 * (size_t)p.nb] = run; run += v3;
  }
  for (; t < t1; ++t) {
    uint32_t* c = p.tile_hist + (size_t)t * p.nb + j;
    const uint32_t v = *c;
    *c = run;
    run += v;
  }
  p.seg_tot[(size_t)seg * p.nb + j] = run;
}

__global__ void __launch_bounds__(256) pt_bucket_tot_kernel(const PtParams p) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.nb) return;
  uint32_t run = 0u;
  for (int s = 0; s < p.segs; ++s) {
    uint32_t* c = p.seg_tot + (size_t)s * p.nb + j;
    const uint32_t v = *c;
    *c = run;
    run += v;
  }
  p.tot[j] = run;
}

struct PtMax {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__global__ void __launch_bounds__(1024) pt_bucket_scan_kernel(const PtParams p) {   // one CTA
  typedef cub::BlockScan<uint32_t, 1024> Scan;
  typedef cub::BlockReduce<uint32_t, 1024> Reduce;
  __shared__ typename Scan::TempStorage s_scan;
  __shared__ typename Reduce::TempStorage s_red;
  const int per = (p.nb + 1023) / 1024;
  const int j0 = threadIdx.x * per;
  uint32_t sum = 0u, mx = 0u;
  for (int k = 0; k < per; ++k)
    if (j0 + k < p.nb) {
      const uint32_t v = p.tot[j0 + k];
      sum += v;
      mx = v > mx ? v : mx;
    }
  uint32_t excl = 0u;
  Scan(s_scan).ExclusiveSum(sum, excl);
  for (int k = 0; k < per; ++k)
    if (j0 + k < p.nb) {
      p.base[j0 + k] = excl;
      excl += p.tot[j0 + k];
    }
  const uint32_t m = Reduce(s_red).Reduce(mx, PtMax());
  if (threadIdx.x == 0) p.stats[1] = m;
}

// ---- scatter: packed pairs grouped by bucket (any order inside a bucket; the bucket kernel sorts) --------------------
__global__ void __launch_bounds__(PT_THREADS) pt_scatter_kernel(const PtParams p) {
  extern __shared__ uint32_t s_dyn[];
  uint32_t* s_cur = s_dyn;
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, seg = tile / p.tiles_per_seg;
  const uint32_t* row = p.tile_hist + (size_t)tile * p.nb;
  const uint32_t* sg = p.seg_tot + (size_t)seg * p.nb;
  for (int j = tid; j < p.nb; j += PT_THREADS) s_cur[j] = p.base[j] + sg[j] + row[j];
  __syncthreads();
  const int64_t i0 = (int64_t)tile * PT_TILE;
  const uint32_t low_mask = (1u << p.low_bits) - 1u;
  const uint32_t minor_mask = (1u << p.minor_bits) - 1u;   // minor_bits <= 31
  bool bad = false;
#pragma unroll 1
  for (int u0 = 0; u0 < PT_ITEMS; u0 += PT_GROUP) {
    int64_t a[PT_GROUP], b[PT_GROUP];
#pragma unroll
    for (int u = 0; u < PT_GROUP; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * PT_THREADS + tid;
      a[u] = i < p.E ? __ldg(p.major + i) : 0;
      b[u] = i < p.E ? __ldg(p.minor + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < PT_GROUP; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * PT_THREADS + tid;
      if (i >= p.E) continue;
      if ((uint64_t)a[u] >= (uint64_t)p.n_major) continue;            // flagged and left out by the count kernel
      if ((uint64_t)b[u] >= (uint64_t)p.n_minor) bad = true;          // keeps its (masked) slot, the call fails
      const uint32_t key = (((uint32_t)a[u] & low_mask) << p.minor_bits) | ((uint32_t)b[u] & minor_mask);
      const uint32_t at = atomicAdd(&s_cur[(uint32_t)(a[u] >> p.low_bits)], 1u);
      p.pairs[at] = make_uint2(key, (uint32_t)i);
    }
  }
  if (bad) atomicOr(p.stats, DEV_ERR_INDEX);
}

// ---- bitonic network on a shared-memory array of any length (every comparator puts the larger value at the higher
// index, so the missing tail of the power-of-two network behaves like +infinity and its comparators are skipped) -----
template <bool BLOCK>
__device__ __forceinline__ void pt_bitonic_shared(uint64_t* a, uint32_t len, int tid, int nthreads) {
  int lp = 0;                                      // log2 of the network's size
  while ((1u << lp) < len) ++lp;
  const uint32_t half = (1u << lp) >> 1;           // comparators per step
  for (int lk = 1; lk <= lp; ++lk) {
    const uint32_t hk = 1u << (lk - 1);
    for (uint32_t t = tid; t < half; t += nthreads) {   // mirror step: r-th of a 2^lk block against its r-th from the end
      const uint32_t first = (t >> (lk - 1)) << lk, r = t & (hk - 1u);
      const uint32_t lo = first + r, hi = first + (2u * hk - 1u - r);
      if (hi < len) {
        const uint64_t x = a[lo], y = a[hi];
        if (x > y) { a[lo] = y; a[hi] = x; }
      }
    }
    if (BLOCK) __syncthreads(); else __syncwarp();
    for (int ld = lk - 2; ld >= 0; --ld) {         // half-cleaners at distance 2^ld
      const uint32_t d = 1u << ld;
      for (uint32_t t = tid; t < half; t += nthreads) {
        const uint32_t lo = ((t >> ld) << (ld + 1)) + (t & (d - 1u)), hi = lo + d;
        if (hi < len) {
          const uint64_t x = a[lo], y = a[hi];
          if (x > y) { a[lo] = y; a[hi] = x; }
        }
      }
      if (BLOCK) __syncthreads(); else __syncwarp();
    }
  }
}

// ---- the same network for a column of up to 32 R entries held in registers: entry r * 32 + lane lives in v[r] of the
// lane.  Distances below 32 are shuffles inside a register, distances of 32 and more pair two registers of the same
// lane (no shuffle), and the mirror step of a block wider than a warp pairs register r with r ^ (block / 32 - 1) of lane
// ^ 31.  Every loop bound is a compile-time constant, so v[] stays in registers. -------------------------------------------
__device__ __forceinline__ void pt_keep(uint64_t& v, uint64_t o, bool keep_min) {
  v = keep_min ? (v < o ? v : o) : (v > o ? v : o);
}

template <int R, int LP>   // a network of 2^LP entries: LP = 5 + log2 R, or fewer stages for a column of <= 2^LP <= 32
__device__ __forceinline__ void pt_bitonic_regs(uint64_t (&v)[R], int lane) {
#pragma unroll
  for (int lk = 1; lk <= LP; ++lk) {
    if (lk <= 5) {
      const int m = (1 << lk) - 1;
      const bool low_side = (lane & (1 << (lk - 1))) == 0;
#pragma unroll
      for (int r = 0; r < R; ++r) pt_keep(v[r], __shfl_xor_sync(0xffffffffu, v[r], m), low_side);
    } else {
      const int mr = (1 << (lk - 5)) - 1, top = 1 << (lk - 6);
      uint64_t o[R];
#pragma unroll
      for (int r = 0; r < R; ++r) o[r] = __shfl_xor_sync(0xffffffffu, v[(r ^ mr) & (R - 1)], 31);
#pragma unroll
      for (int r = 0; r < R; ++r) pt_keep(v[r], o[r], (r & top) == 0);
    }
#pragma unroll
    for (int ld = lk - 2; ld >= 0; --ld) {
      if (ld >= 5) {
        const int dr = 1 << (ld - 5);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if ((r & dr) == 0) {
            const uint64_t a = v[r], b = v[(r ^ dr) & (R - 1)];
            v[r] = a < b ? a : b;
            v[(r ^ dr) & (R - 1)] = a < b ? b : a;
          }
      } else {
        const int m = 1 << ld;
        const bool low_side = (lane & m) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) pt_keep(v[r], __shfl_xor_sync(0xffffffffu, v[r], m), low_side);
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void pt_sort_column_regs(const uint64_t* col, uint32_t len, int lane, int64_t* out_i,
                                                    int64_t* out_p) {
  uint64_t v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = (uint32_t)(r * 32 + lane) < len ? col[r * 32 + lane] : ~0ull;
  pt_bitonic_regs<R, 5 + (R == 2 ? 1 : R == 4 ? 2 : 3)>(v, lane);
#pragma unroll
  for (int r = 0; r < R; ++r)
    if ((uint32_t)(r * 32 + lane) < len) {
      out_i[r * 32 + lane] = (int64_t)(v[r] >> 32);
      out_p[r * 32 + lane] = (int64_t)(v[r] & 0xffffffffull);
    }
}

// ---- one CTA per bucket: counting sort by major id in shared memory, then every column sorted by minor id ------------
__global__ void __launch_bounds__(PT_BTHREADS, 2) pt_bucket_kernel(const PtParams p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  typedef cub::BlockScan<uint32_t, PT_BTHREADS> Scan;
  __shared__ typename Scan::TempStorage s_scan;
  __shared__ int s_heavy[PT_HEAVY_MAX];
  __shared__ int s_nheavy;
  const int ncol = 1 << p.low_bits;
  uint64_t* s_pair = reinterpret_cast<uint64_t*>(s_raw);        // [PT_CAP] minor << 32 | edge id, grouped by column
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_pair + PT_CAP);   // [ncol] entries of the column
  uint32_t* s_off = s_cnt + ncol;                               // [ncol] first slot of the column inside the bucket
  uint32_t* s_cur = s_off + ncol;                               // [ncol] cursor of the grouping pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bkt = blockIdx.x;
  const uint32_t n = p.tot[bkt], start = p.base[bkt];
  for (int c = tid; c < ncol; c += PT_BTHREADS) s_cnt[c] = 0u;
  if (tid == 0) s_nheavy = 0;
  __syncthreads();
  const uint2* src = p.pairs + start;
  for (uint32_t i = tid; i < n; i += 4 * PT_BTHREADS) {        // four keys in flight per thread (first read: DRAM)
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) k[u] = i + u * PT_BTHREADS < n ? src[i + u * PT_BTHREADS].x : 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * PT_BTHREADS < n) atomicAdd(&s_cnt[k[u] >> p.minor_bits], 1u);
  }
  __syncthreads();
  {
    const int per = (ncol + PT_BTHREADS - 1) / PT_BTHREADS;    // 1 or 2 columns per thread
    const int c0 = tid * per;
    uint32_t v[2] = {0u, 0u};
    for (int k = 0; k < per; ++k)
      if (c0 + k < ncol) v[k] = s_cnt[c0 + k];
    uint32_t excl = 0u;
    Scan(s_scan).ExclusiveSum(v[0] + v[1], excl);
    for (int k = 0; k < per; ++k) {
      const int c = c0 + k;
      if (c >= ncol) break;
      s_off[c] = excl;
      s_cur[c] = excl;
      const int64_t col = ((int64_t)bkt << p.low_bits) + c;
      if (col < p.n_major) p.ptrs[col] = (int64_t)start + excl;           // ind2ptr (storage.rs:67-101)
      if (v[k] > (uint32_t)PT_WARP_MAX) s_heavy[atomicAdd(&s_nheavy, 1)] = c;
      excl += v[k];
    }
    if (bkt == p.nb - 1 && tid == 0) p.ptrs[p.n_major] = p.E;
  }
  __syncthreads();
  const uint32_t minor_mask = (1u << p.minor_bits) - 1u;
  for (uint32_t i = tid; i < n; i += 4 * PT_BTHREADS) {        // second read of the bucket: served by the L2
    uint2 kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) kv[u] = i + u * PT_BTHREADS < n ? src[i + u * PT_BTHREADS] : make_uint2(0u, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * PT_BTHREADS < n) {
        const uint32_t at = atomicAdd(&s_cur[kv[u].x >> p.minor_bits], 1u);
        s_pair[at] = ((uint64_t)(kv[u].x & minor_mask) << 32) | kv[u].y;
      }
  }
  __syncthreads();
  int64_t* out_i = p.indices + start;
  int64_t* out_p = p.perm + start;
  for (int c = warp; c < ncol; c += PT_BTHREADS / 32) {
    const uint32_t len = s_cnt[c], off = s_off[c];
    if (len == 0u || len > (uint32_t)PT_WARP_MAX) continue;
    if (len <= 32u) {   // the column in registers, one entry per lane; the network no wider than the column needs
      uint64_t v[1] = {(uint32_t)lane < len ? s_pair[off + lane] : ~0ull};
      if (len > 16u) pt_bitonic_regs<1, 5>(v, lane);
      else if (len > 8u) pt_bitonic_regs<1, 4>(v, lane);
      else if (len > 4u) pt_bitonic_regs<1, 3>(v, lane);
      else if (len > 2u) pt_bitonic_regs<1, 2>(v, lane);
      else if (len > 1u) pt_bitonic_regs<1, 1>(v, lane);
      if ((uint32_t)lane < len) {
        out_i[off + lane] = (int64_t)(v[0] >> 32);
        out_p[off + lane] = (int64_t)(v[0] & 0xffffffffull);
      }
    } else if (len <= 64u) {   // two entries per lane.  (Four and eight per lane, for columns of up to 128 and 256 entries,
                               // measured SLOWER than the shared-memory network below: 2.06 -> 2.47 ms for this kernel on the
                               // lognormal products-shaped graph, whose columns of 65..256 entries hold 29 % of the edges.)
      pt_sort_column_regs<2>(s_pair + off, len, lane, out_i + off, out_p + off);
    } else {
      pt_bitonic_shared<false>(s_pair + off, len, lane, 32);
      for (uint32_t i = lane; i < len; i += 32) {
        const uint64_t v = s_pair[off + i];
        out_i[off + i] = (int64_t)(v >> 32);
        out_p[off + i] = (int64_t)(v & 0xffffffffull);
      }
    }
  }
  __syncthreads();
  const int nheavy = s_nheavy;
  for (int h = 0; h < nheavy; ++h) {   // the few columns above PT_WARP_MAX entries: the whole CTA sorts each
    const int c = s_heavy[h];
    const uint32_t len = s_cnt[c], off = s_off[c];
    pt_bitonic_shared<true>(s_pair + off, len, tid, PT_BTHREADS);
    for (uint32_t i = tid; i < len; i += PT_BTHREADS) {
      const uint64_t v = s_pair[off + i];
      out_i[off + i] = (int64_t)(v >> 32);
      out_p[off + i] = (int64_t)(v & 0xffffffffull);
    }
  }
}

struct PtLayout {
  int low_bits, nb, tiles, tiles_per_seg, segs;
  size_t off_stats, off_pairs, off_tile_hist, off_seg_tot, off_tot, off_base, total;
};

// bytes the partition form can need for E edges, whatever the shape (the workspace query does not know which side is major)
inline size_t pt_workspace_bound(int64_t E) {
  const size_t n = (size_t)(E > 0 ? E : 1);
  return 256 + align_up(n * 8) + align_up(n * 4 + (size_t)PT_NB_MAX * 4) + align_up((size_t)PT_SEGS * PT_NB_MAX * 4) +
         2 * align_up((size_t)PT_NB_MAX * 4) + 256;
}

// Whether the partition form applies, and its geometry: buckets of 2^low_bits major ids that hold about half of PT_CAP on
// average (the largest bucket is checked on the device before the scatter), key = low bits + minor id in 32 bits, the
// histogram matrix no larger than 4 bytes per edge.
inline bool pt_plan(int64_t E, int64_t n_major, int minor_bits, PtLayout& L) {
  if (E <= 0 || n_major <= 0 || minor_bits > 31) return false;
  const double avg = (double)E / (double)n_major;
  int low_bits = 0;
  while (low_bits < PT_MAX_LOW_BITS && low_bits + 1 + minor_bits <= 32 && avg * (double)(2ll << low_bits) <= 0.6 * PT_CAP)
    ++low_bits;
  if (avg * (double)(1ll << low_bits) > 0.6 * PT_CAP) return false;   // one major id alone is a bucket's worth
  const int64_t nb = (n_major + ((int64_t)1 << low_bits) - 1) >> low_bits;
  if (nb > PT_NB_MAX) return false;
  const int64_t tiles = (E + PT_TILE - 1) / PT_TILE;
  if ((size_t)tiles * (size_t)nb > (size_t)E + (size_t)PT_NB_MAX) return false;
  L.low_bits = low_bits;
  L.nb = (int)nb;
  L.tiles = (int)tiles;
  L.tiles_per_seg = (int)((tiles + PT_SEGS - 1) / PT_SEGS);
  L.segs = (int)((tiles + L.tiles_per_seg - 1) / L.tiles_per_seg);
  L.off_stats = 0;
  L.off_pairs = 256;
  L.off_tile_hist = L.off_pairs + align_up((size_t)E * 8);
  L.off_seg_tot = L.off_tile_hist + align_up((size_t)tiles * nb * 4);
  L.off_tot = L.off_seg_tot + align_up((size_t)L.segs * nb * 4);
  L.off_base = L.off_tot + align_up((size_t)nb * 4);
  L.total = L.off_base + align_up((size_t)nb * 4) + 256;
  return true;
}

inline bool pt_wanted() {   // TCHGEO_CSX_SORT=partition | cub
  const char* v = getenv("TCHGEO_CSX_SORT");
  if (v == nullptr || *v == 0) return PT_DEFAULT_ON;
  return strcmp(v, "partition") == 0;
}

// -> TCHGEO_OK with *done = true when the outputs are written, *done = false when the radix form has to run instead
tchgeo_status csx_partition_form(const int64_t* major, const int64_t* minor, int64_t E, int64_t n_major, int64_t n_minor,
                                 int minor_bits, int64_t* ptrs, int64_t* indices, int64_t* perm, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream, bool* done) {
  *done = false;
  PtLayout L;
  if (!pt_plan(E, n_major, minor_bits, L) || workspace_bytes < L.total) return TCHGEO_OK;
  const int ncol = 1 << L.low_bits;
  const size_t hist_smem = (size_t)L.nb * 4;
  const size_t bucket_smem = (size_t)PT_CAP * 8 + (size_t)ncol * 12;
  static bool configured[64] = {};   // per device; benign race: the attributes are idempotent
  int dev = 0;
  TCHGEO_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(pt_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_NB_MAX * 4));
    TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(pt_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_NB_MAX * 4));
    TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(pt_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           PT_CAP * 8 + (1 << PT_MAX_LOW_BITS) * 12));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  char* ws = (char*)workspace;
  PtParams p;
  p.major = major; p.minor = minor; p.E = E; p.n_major = n_major; p.n_minor = n_minor;
  p.low_bits = L.low_bits; p.minor_bits = minor_bits; p.nb = L.nb;
  p.tiles = L.tiles; p.tiles_per_seg = L.tiles_per_seg; p.segs = L.segs;
  p.stats = (uint32_t*)(ws + L.off_stats);
  p.pairs = (uint2*)(ws + L.off_pairs);
  p.tile_hist = (uint32_t*)(ws + L.off_tile_hist);
  p.seg_tot = (uint32_t*)(ws + L.off_seg_tot);
  p.tot = (uint32_t*)(ws + L.off_tot);
  p.base = (uint32_t*)(ws + L.off_base);
  p.ptrs = ptrs; p.indices = indices; p.perm = perm;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(p.stats, 0, 256, stream));
  pt_count_kernel<<<(unsigned)L.tiles, PT_THREADS, hist_smem, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  const unsigned gb = (unsigned)((L.nb + 255) / 256);
  pt_seg_prefix_kernel<<<dim3(gb, (unsigned)L.segs), 256, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  pt_bucket_tot_kernel<<<gb, 256, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  pt_bucket_scan_kernel<<<1, 1024, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t hstats[2] = {0u, 0u};
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(hstats, p.stats, 8, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (hstats[0] != 0u) {
    *done = true;
    return status_from_dev_err(hstats[0]);
  }
  if (hstats[1] > (uint32_t)PT_CAP) return TCHGEO_OK;   // a bucket does not fit one CTA's shared memory: radix form
  pt_scatter_kernel<<<(unsigned)L.tiles, PT_THREADS, hist_smem, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  pt_bucket_kernel<<<(unsigned)L.nb, PT_BTHREADS, bucket_smem, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t herr = 0u;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, p.stats, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  *done = true;
  return status_from_dev_err(herr);
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_ind2ptr(const int64_t* ind, int64_t numel, int64_t m, int64_t* out,
                                        tchgeo_stream stream_) {
  TCHGEO_REQUIRE(numel >= 0 && m >= 0 && out != nullptr && (numel == 0 || ind != nullptr), "bad ind2ptr argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  ind2ptr_kernel<<<grid_for(numel > 0 ? numel : m + 1), CSX_THREADS, 0, stream>>>(ind, numel, m, out);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}

extern "C" size_t tchgeo_coo_to_csx_workspace_bytes(int64_t num_edges, int64_t n_rows, int64_t n_cols) {
  if (num_edges < 0 || n_rows < 0 || n_cols < 0) return 0;
  CsxLayout L;
  if (csx_layout(num_edges, 64, L) != cudaSuccess) return 0;
  const size_t pt = pt_workspace_bound(num_edges);
  return L.total > pt ? L.total : pt;
}

extern "C" tchgeo_status tchgeo_coo_to_csx(const int64_t* row, const int64_t* col, int64_t E, int64_t n_rows,
                                           int64_t n_cols, int32_t csc, int64_t* ptrs, int64_t* indices, int64_t* perm,
                                           void* workspace, size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(E >= 0 && n_rows >= 0 && n_cols >= 0, "negative size");
  TCHGEO_REQUIRE(E < ((int64_t)1 << 32), "more than 2^32-1 edges are not supported");
  TCHGEO_REQUIRE(ptrs != nullptr && (E == 0 || (row && col && indices && perm)), "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n_major = csc ? n_cols : n_rows;
  const int64_t n_minor = csc ? n_rows : n_cols;
  const int64_t* major = csc ? col : row;
  const int64_t* minor = csc ? row : col;
  if (E == 0) {  // storage.rs:78-80
    fill_zero_i64_kernel<<<grid_for(n_major + 1), CSX_THREADS, 0, stream>>>(ptrs, n_major + 1);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    return TCHGEO_OK;
  }
  const int minor_bits = bits_for(n_minor), major_bits = bits_for(n_major);
  TCHGEO_REQUIRE(minor_bits + major_bits <= 64, "graph too large: (row, col) key needs more than 64 bits");
  const int end_bit = minor_bits + major_bits;
  TCHGEO_REQUIRE(workspace != nullptr, "workspace is NULL");
  if (pt_wanted()) {
    bool done = false;
    const tchgeo_status st = csx_partition_form(major, minor, E, n_major, n_minor, minor_bits, ptrs, indices, perm,
                                                workspace, workspace_bytes, stream, &done);
    if (st != TCHGEO_OK || done) return st;
  }
  CsxLayout L;
  TCHGEO_CUDA_CHECK(csx_layout(E, end_bit, L));
  TCHGEO_REQUIRE(workspace != nullptr, "workspace is NULL");
  if (workspace_bytes < L.total) {
    set_last_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return TCHGEO_ERR_CAPACITY;
  }
  char* ws = (char*)workspace;
  uint32_t* err = (uint32_t*)(ws + L.off_err);
  uint64_t* keys0 = (uint64_t*)(ws + L.off_keys0);
  uint64_t* keys1 = (uint64_t*)(ws + L.off_keys1);
  uint32_t* vals0 = (uint32_t*)(ws + L.off_vals0);
  uint32_t* vals1 = (uint32_t*)(ws + L.off_vals1);
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(err, 0, 256, stream));
  build_keys_kernel<<<grid_for(E), CSX_THREADS, 0, stream>>>(major, minor, E, n_major, n_minor, minor_bits, keys0,
                                                             vals0, err);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  cub::DoubleBuffer<uint64_t> dk(keys0, keys1);
  cub::DoubleBuffer<uint32_t> dv(vals0, vals1);
  size_t cub_bytes = L.cub_bytes;
  TCHGEO_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.off_cub, cub_bytes, dk, dv, E, 0, end_bit, stream));
  decode_kernel<<<grid_for(E), CSX_THREADS, 0, stream>>>(dk.Current(), dv.Current(), E, n_major, minor_bits, ptrs,
                                                         indices, perm);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t herr = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(herr);
}
