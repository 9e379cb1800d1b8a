// Compact host transport of sampled batches (extension; the reference builds its output tensors in host memory to
// begin with, src/python.rs:259-262).
//
// The end-to-end path of a step is bound by PCIe: the reference layout is three i64 vectors per sampled edge (samples,
// cols, edge_index; rows is an arange) -- 24 B per edge, 3.7 GB per headline step at ~56 GB/s.  Two of the three carry
// values below 2^31 and the third is a run-length sequence, so a group of batches travels as
//     samples32 [sum n_b] i32 | eidx32 [sum e_b] i32 | counts [sum n_b] u8 (edges drawn for node j = run length of j in cols)
// = 9 B per edge, and tchgeo_host_unpack_transport (host_unpack.cpp) rebuilds the reference's i64 vectors in host memory
// with a few threads and non-temporal stores while the next group is still on the bus.  The result is the same bytes the
// plain copy lands.
#include "common.cuh"

namespace tchgeo {
namespace {

constexpr int TP_THREADS = 256;
constexpr int TP_ITEMS = 8;
constexpr int TP_CHUNK = TP_THREADS * TP_ITEMS;

// offsets[0..B] of the clamped lengths (one CTA)
__global__ void __launch_bounds__(1024) tp_offsets_kernel(const int64_t* __restrict__ lens, int64_t B, int64_t max_len,
                                                          int64_t* __restrict__ off) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < B; b0 += 1024) {
    const int64_t b = b0 + tid;
    int64_t v = b < B ? lens[b] : 0;
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

// used prefix of every padded row -> one packed i32 vector
__global__ void __launch_bounds__(TP_THREADS) tp_pack32_kernel(const int64_t* __restrict__ src, int64_t stride,
                                                              const int64_t* __restrict__ off, int32_t* __restrict__ dst,
                                                              uint32_t* err) {
  const int64_t b = blockIdx.y;
  const int64_t n = off[b + 1] - off[b];
  const int64_t i0 = (int64_t)blockIdx.x * TP_CHUNK;
  if (i0 >= n) return;
  const int64_t* s = src + b * stride + i0;
  int32_t* d = dst + off[b] + i0;
  const int64_t m = min((int64_t)TP_CHUNK, n - i0);
  int64_t v[TP_ITEMS];
#pragma unroll
  for (int u = 0; u < TP_ITEMS; ++u) {
    const int64_t i = u * TP_THREADS + threadIdx.x;
    v[u] = i < m ? __ldcs(s + i) : 0;
  }
  bool bad = false;
#pragma unroll
  for (int u = 0; u < TP_ITEMS; ++u) {
    const int64_t i = u * TP_THREADS + threadIdx.x;
    bad |= (v[u] < 0 || v[u] > 0x7fffffffll);
    if (i < m) __stcs(d + i, (int32_t)v[u]);
  }
  if (bad) atomicOr(err, DEV_ERR_INDEX);
}

// counts[n_off[b] + j] = length of the run of j in the batch's cols (cols is non-decreasing: edges are emitted frontier
// node by frontier node).  The thread that sees a run start measures it; counts is zeroed beforehand.
__global__ void __launch_bounds__(TP_THREADS) tp_runs_kernel(const int64_t* __restrict__ cols, int64_t stride,
                                                            const int64_t* __restrict__ e_off,
                                                            const int64_t* __restrict__ n_off, uint8_t* __restrict__ counts,
                                                            uint32_t* err) {
  const int64_t b = blockIdx.y;
  const int64_t ne = e_off[b + 1] - e_off[b], nn = n_off[b + 1] - n_off[b];
  const int64_t i0 = (int64_t)blockIdx.x * TP_CHUNK;
  if (i0 >= ne) return;
  const int64_t* c = cols + b * stride;
  uint8_t* out = counts + n_off[b];
  bool bad = false;
#pragma unroll 2
  for (int u = 0; u < TP_ITEMS; ++u) {
    const int64_t e = i0 + u * TP_THREADS + threadIdx.x;
    if (e >= ne) continue;
    const int64_t cur = __ldg(c + e);
    const int64_t prev = e > 0 ? __ldg(c + e - 1) : -1;
    if (cur == prev) continue;
    if (cur < prev || cur >= nn) { bad = true; continue; }   // not the layout this transport is defined for
    int len = 1;
    while (e + len < ne && len <= 255 && __ldg(c + e + len) == cur) ++len;
    if (len > 255) { bad = true; continue; }
    out[cur] = (uint8_t)len;
  }
  if (bad) atomicOr(err, DEV_ERR_CAPACITY);
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_pack_transport(const int64_t* samples, int64_t samples_stride, const int64_t* cols,
                                               const int64_t* edge_index, int64_t edges_stride, const int64_t* n_lens,
                                               const int64_t* e_lens, int64_t count, int64_t max_n, int64_t max_e,
                                               int32_t* samples32, int32_t* eidx32, uint8_t* counts, int64_t counts_bytes,
                                               int64_t* n_off, int64_t* e_off, int32_t* err_word, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(count >= 0 && count <= 65535 && samples_stride >= 0 && edges_stride >= 0 && max_n >= 0 &&
                     max_n <= samples_stride && max_e >= 0 && max_e <= edges_stride && counts_bytes >= 0,
                 "bad transport geometry");
  if (count == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(n_lens && e_lens && n_off && e_off && err_word && samples32 && counts, "NULL pointer");
  TCHGEO_REQUIRE((max_n == 0 || samples) && (max_e == 0 || (cols && (edge_index || !eidx32))), "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  tp_offsets_kernel<<<1, 1024, 0, stream>>>(n_lens, count, max_n, n_off);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  tp_offsets_kernel<<<1, 1024, 0, stream>>>(e_lens, count, max_e, e_off);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  if (counts_bytes > 0) TCHGEO_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)counts_bytes, stream));
  uint32_t* err = (uint32_t*)err_word;
  if (max_n > 0) {
    const dim3 grid((unsigned)((max_n + TP_CHUNK - 1) / TP_CHUNK), (unsigned)count);
    tp_pack32_kernel<<<grid, TP_THREADS, 0, stream>>>(samples, samples_stride, n_off, samples32, err);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  if (max_e > 0) {
    const dim3 grid((unsigned)((max_e + TP_CHUNK - 1) / TP_CHUNK), (unsigned)count);
    if (eidx32) {   // NULL: edge_index travels as it is (tchgeo_pack_ragged), the host then has less to rebuild
      tp_pack32_kernel<<<grid, TP_THREADS, 0, stream>>>(edge_index, edges_stride, e_off, eidx32, err);
      TCHGEO_CUDA_CHECK(cudaGetLastError());
    }
    tp_runs_kernel<<<grid, TP_THREADS, 0, stream>>>(cols, edges_stride, e_off, n_off, counts, err);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  return TCHGEO_OK;
}
