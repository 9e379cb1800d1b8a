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
