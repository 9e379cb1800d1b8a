// Row gather on sm_100a: dst[i, :] = src[index[i], :].
//
// The step that follows the sampler in every loader built on the reference (examples/neighbor_sampling.py:21-24,
// PyG's filter_data: x[samples], edge_attr[perm[edge_index]]).  Pure HBM traffic: 8 B of index, row_bytes read and
// row_bytes written per gathered row, so the kernel is organised around full-width memory transactions:
// rows are cut into 16-byte vectors (8 / 4 / 1 bytes when the row size or the base pointers are not 16-byte
// aligned), consecutive threads take consecutive vectors of the same row, and every thread keeps four independent
// vectors in flight.  Reads of a row are contiguous (a 400-byte feature row is 3-4 DRAM lines), writes are
// streaming stores that bypass L2 residency.  No tensor cores: there is no arithmetic.
#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int GT_THREADS = 256;
constexpr int GT_UNROLL = 4;

template <typename V>
__device__ __forceinline__ V ld_row(const V* p) { return __ldg(p); }

template <typename V>
__device__ __forceinline__ void st_stream(V* p, const V& v) { __stcs(p, v); }

template <typename V>
__global__ void __launch_bounds__(GT_THREADS) gather_rows_kernel(const V* __restrict__ src, int64_t num_rows,
                                                                int64_t vecs_per_row, const int64_t* __restrict__ index,
                                                                int64_t n, V* __restrict__ dst, uint32_t* err) {
  // vector t of the output is (row i = t / vecs_per_row, column c = t % vecs_per_row); the pair is divided out once
  // per thread and then advanced by the grid stride with adds only (a 64-bit division per vector would cost more
  // issue slots than the copy itself)
  const int64_t stride = (int64_t)gridDim.x * GT_THREADS;
  const int64_t di = stride / vecs_per_row, dc = stride - di * vecs_per_row;
  const int64_t t0 = (int64_t)blockIdx.x * GT_THREADS + threadIdx.x;
  int64_t i = t0 / vecs_per_row, c = t0 - i * vecs_per_row;
  while (i < n) {
    V v[GT_UNROLL];
    int64_t o[GT_UNROLL];
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u) {
      o[u] = -1;
      if (i < n) {
        const int64_t r = __ldg(index + i);
        if (r < 0 || r >= num_rows) {
          if (c == 0 && err) atomicOr(err, DEV_ERR_INDEX);
        } else {
          v[u] = ld_row(src + r * vecs_per_row + c);
          o[u] = i * vecs_per_row + c;
        }
      }
      i += di;
      c += dc;
      if (c >= vecs_per_row) { c -= vecs_per_row; ++i; }
    }
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u)
      if (o[u] >= 0) st_stream(dst + o[u], v[u]);
  }
}

template <typename V>
cudaError_t launch_gather(const void* src, int64_t num_rows, int64_t row_bytes, const int64_t* index, int64_t n, void* dst,
                          uint32_t* err, cudaStream_t stream) {
  const int64_t vpr = row_bytes / (int64_t)sizeof(V);
  const int64_t total = n * vpr;
  int64_t grid = (total + (int64_t)GT_THREADS * GT_UNROLL - 1) / ((int64_t)GT_THREADS * GT_UNROLL);
  const int64_t cap = 148 * 8 * 16;  // grid-stride beyond 16 waves of 8 resident CTAs per SM
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  gather_rows_kernel<V><<<(unsigned)grid, GT_THREADS, 0, stream>>>((const V*)src, num_rows, vpr, index, n, (V*)dst, err);
  return cudaGetLastError();
}

// ---- ragged pack: the used prefixes of B padded rows, back to back -------------------------------------------
// dst[off[b] + i] = src[b * stride + i] for i < lens[b], off = exclusive prefix sum of lens.  What a host-side consumer
// wants before a D2H copy: the sampler's outputs live in padded [B, capacity] buffers (about 60 % used), and one copy of
// the packed prefixes moves only the used bytes.  The offsets are computed on the device by one CTA (B is small).
constexpr int PK_THREADS = 256;
constexpr int PK_CHUNK = PK_THREADS * 8;  // elements per CTA

__global__ void __launch_bounds__(1024) pack_offsets_kernel(const int64_t* __restrict__ lens, int64_t lens_stride, int64_t B,
                                                            int64_t max_len, int64_t* __restrict__ off) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < B; b0 += 1024) {
    const int64_t b = b0 + tid;
    int64_t v = b < B ? lens[b * lens_stride] : 0;
    v = v < 0 ? 0 : (v > max_len ? max_len : v);
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (b < B) off[b] = before + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) off[B] = s_carry;
}

__global__ void __launch_bounds__(PK_THREADS) pack_kernel(const int64_t* __restrict__ src, int64_t stride,
                                                         const int64_t* __restrict__ off, int64_t* __restrict__ dst) {
  const int64_t b = blockIdx.y;
  const int64_t n = off[b + 1] - off[b];
  const int64_t i0 = (int64_t)blockIdx.x * PK_CHUNK;
  if (i0 >= n) return;
  const int64_t* s = src + b * stride + i0;
  int64_t* d = dst + off[b] + i0;
  const int64_t m = min((int64_t)PK_CHUNK, n - i0);
  int64_t v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t i = u * PK_THREADS + threadIdx.x;
    v[u] = i < m ? __ldcs(s + i) : 0;
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t i = u * PK_THREADS + threadIdx.x;
    if (i < m) __stcs(d + i, v[u]);
  }
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_gather_rows(const void* src, int64_t num_rows, int64_t row_bytes, const int64_t* index,
                                            int64_t n, void* dst, int32_t* scratch, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_rows >= 0 && row_bytes >= 0 && n >= 0, "negative size");
  if (n == 0 || row_bytes == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(src && index && dst, "NULL pointer");
  TCHGEO_REQUIRE(n <= ((int64_t)1 << 62) / row_bytes, "gather too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (scratch) TCHGEO_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4, stream));
  const uintptr_t align = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)row_bytes;
  cudaError_t e;
  if ((align & 15u) == 0) e = launch_gather<uint4>(src, num_rows, row_bytes, index, n, dst, (uint32_t*)scratch, stream);
  else if ((align & 7u) == 0) e = launch_gather<uint2>(src, num_rows, row_bytes, index, n, dst, (uint32_t*)scratch, stream);
  else if ((align & 3u) == 0) e = launch_gather<uint32_t>(src, num_rows, row_bytes, index, n, dst, (uint32_t*)scratch, stream);
  else e = launch_gather<uint8_t>(src, num_rows, row_bytes, index, n, dst, (uint32_t*)scratch, stream);
  TCHGEO_CUDA_CHECK(e);
  if (!scratch) return TCHGEO_OK;  // asynchronous: no validation read-back (rows with a bad index are left unwritten)
  uint32_t h = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&h, scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(h);
}

extern "C" tchgeo_status tchgeo_pack_ragged(const int64_t* src, int64_t stride, const int64_t* lens, int64_t lens_stride,
                                            int64_t num_batches, int64_t max_len, int64_t* dst, int64_t* offsets,
                                            tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_batches >= 0 && num_batches <= 65535 && stride >= 0 && max_len >= 0 && max_len <= stride && lens_stride >= 1,
                 "bad pack geometry");
  if (num_batches == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(lens && offsets && (max_len == 0 || (src && dst)), "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  pack_offsets_kernel<<<1, 1024, 0, stream>>>(lens, lens_stride, num_batches, max_len, offsets);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  if (max_len == 0) return TCHGEO_OK;
  const dim3 grid((unsigned)((max_len + PK_CHUNK - 1) / PK_CHUNK), (unsigned)num_batches);
  pack_kernel<<<grid, PK_THREADS, 0, stream>>>(src, stride, offsets, dst);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
}
