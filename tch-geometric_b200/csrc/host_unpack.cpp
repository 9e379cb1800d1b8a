// Host half of the compact transport (see transport.cu): samples32 / eidx32 / counts -> the reference's i64 samples,
// edge_index and cols vectors in host memory.  Plain C++ (no CUDA): worker threads take batches from a shared counter;
// every output line is written once with non-temporal stores (the vectors are far larger than the caches and are read
// by somebody else later), 64 bytes at a time where the CPU has AVX-512, 16 bytes otherwise.
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/tchgeo_cuda.h"

namespace tchgeo {
void set_last_error(const char* fmt, ...);
}

namespace {

#if defined(__x86_64__)
inline void store_nt(int64_t* p, int64_t v) { _mm_stream_si64(reinterpret_cast<long long*>(p), (long long)v); }

void widen_sse2(const int32_t* s, int64_t n, int64_t* d) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 15u)) { store_nt(d + i, s[i]); ++i; }
  for (; i + 4 <= n; i += 4) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
    const __m128i sign = _mm_srai_epi32(v, 31);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), _mm_unpacklo_epi32(v, sign));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 2), _mm_unpackhi_epi32(v, sign));
  }
  for (; i < n; ++i) store_nt(d + i, s[i]);
}

__attribute__((target("avx512f"))) void widen_avx512(const int32_t* s, int64_t n, int64_t* d) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 63u)) { store_nt(d + i, s[i]); ++i; }
  for (; i + 8 <= n; i += 8) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
    _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), _mm512_cvtepi32_epi64(v));
  }
  for (; i < n; ++i) store_nt(d + i, s[i]);
}

// i64 values staged in a cache-resident buffer -> destination, whole lines at a time
void flush_sse2(const int64_t* buf, int64_t n, int64_t* d) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 15u)) { store_nt(d + i, buf[i]); ++i; }
  for (; i + 2 <= n; i += 2)
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + i)));
  for (; i < n; ++i) store_nt(d + i, buf[i]);
}

__attribute__((target("avx512f"))) void flush_avx512(const int64_t* buf, int64_t n, int64_t* d) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 63u)) { store_nt(d + i, buf[i]); ++i; }
  for (; i + 8 <= n; i += 8)
    _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), _mm512_loadu_si512(reinterpret_cast<const void*>(buf + i)));
  for (; i < n; ++i) store_nt(d + i, buf[i]);
}

// ordinary (write-allocate) stores: a core keeps more lines in flight with them than through its few write-combining
// buffers, so ONE thread is faster this way (measured: 6.0 against 4.0 GB/s), while non-temporal stores move half the
// DRAM bytes and win once all cores are busy -- TCHGEO_HOST_STORES=regular|nt picks, default nt
__attribute__((target("avx512f"))) void widen_avx512_regular(const int32_t* s, int64_t n, int64_t* d) {
  int64_t i = 0;
  for (; i + 8 <= n; i += 8)
    _mm512_storeu_si512(reinterpret_cast<void*>(d + i),
                        _mm512_cvtepi32_epi64(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i))));
  for (; i < n; ++i) d[i] = s[i];
}
void widen_regular(const int32_t* s, int64_t n, int64_t* d) {
  for (int64_t i = 0; i < n; ++i) d[i] = s[i];
}

std::atomic<int> g_mode(-1);   // bit 0: 64-byte vectors (AVX-512); bit 1: regular instead of non-temporal stores
void choose_simd() {
  const char* e = getenv("TCHGEO_HOST_SIMD");   // "sse2" forces the narrow path (tests)
  const char* st = getenv("TCHGEO_HOST_STORES");
  int m = (e && strcmp(e, "sse2") == 0) ? 0 : (__builtin_cpu_supports("avx512f") ? 1 : 0);
  if (st && strcmp(st, "regular") == 0) m |= 2;
  g_mode.store(m);
}
inline void widen(const int32_t* s, int64_t n, int64_t* d) {
  switch (g_mode.load(std::memory_order_relaxed)) {
    case 1: widen_avx512(s, n, d); break;
    case 2: widen_regular(s, n, d); break;
    case 3: widen_avx512_regular(s, n, d); break;
    default: widen_sse2(s, n, d); break;
  }
}
inline void flush(const int64_t* b, int64_t n, int64_t* d) {
  switch (g_mode.load(std::memory_order_relaxed)) {
    case 1: flush_avx512(b, n, d); break;
    case 2: case 3: memcpy(d, b, (size_t)n * 8); break;
    default: flush_sse2(b, n, d); break;
  }
}
inline void fence() { _mm_sfence(); }
#else
inline void widen(const int32_t* s, int64_t n, int64_t* d) { for (int64_t i = 0; i < n; ++i) d[i] = s[i]; }
inline void flush(const int64_t* b, int64_t n, int64_t* d) { memcpy(d, b, (size_t)n * 8); }
inline void fence() {}
inline void choose_simd() {}
#endif

// cols = 0 x counts[0], 1 x counts[1], ...; false when the runs do not add up to n_edges.  The values are produced in a
// cache-resident buffer and leave it in pieces that END on a 64-byte boundary of the destination, so every line of the
// destination is written by exactly one run of full-width non-temporal stores.
bool expand_runs(const uint8_t* counts, int64_t n_nodes, int64_t* d, int64_t n_edges) {
  constexpr int BUF = 1024;                  // 8 KB: stays in the L1
  alignas(64) int64_t buf[BUF + 256];
  int64_t e = 0;                             // edges flushed so far
  int fill = 0;
  for (int64_t j = 0; j < n_nodes; ++j) {
    const int c = counts[j];
    for (int k = 0; k < c; ++k) buf[fill + k] = j;
    fill += c;
    if (fill >= BUF) {
      if (e + fill > n_edges) return false;
      const int keep = (int)((reinterpret_cast<uintptr_t>(d + e + fill) & 63u) >> 3);   // the piece ends line-aligned
      const int m = fill - keep;
      flush(buf, m, d + e);
      e += m;
      memmove(buf, buf + m, (size_t)keep * 8);
      fill = keep;
    }
  }
  if (e + fill != n_edges) return false;
  flush(buf, fill, d + e);
  return true;
}

}  // namespace

extern "C" TCHGEO_API tchgeo_status tchgeo_host_unpack_transport(const int32_t* samples32, const int32_t* eidx32,
                                                                 const uint8_t* counts, const int64_t* n_off,
                                                                 const int64_t* e_off, int64_t count, int64_t* samples,
                                                                 int64_t* cols, int64_t* edge_index, int32_t num_threads) {
  if (!(count >= 0 && num_threads >= 1 && num_threads <= 1024)) {
    tchgeo::set_last_error("bad unpack argument");
    return TCHGEO_ERR_BAD_ARG;
  }
  if (count == 0) return TCHGEO_OK;
  if (!(samples32 && counts && n_off && e_off && samples && cols && ((eidx32 != nullptr) == (edge_index != nullptr)))) {
    tchgeo::set_last_error("NULL pointer");
    return TCHGEO_ERR_BAD_ARG;
  }
  for (int64_t b = 0; b < count; ++b)
    if (!(n_off[b] <= n_off[b + 1] && e_off[b] <= e_off[b + 1] && n_off[0] >= 0 && e_off[0] >= 0)) {
      tchgeo::set_last_error("offsets must be non-decreasing");
      return TCHGEO_ERR_BAD_ARG;
    }
  choose_simd();
  // work items: (batch, vector) -- three times as many as batches, so that a few threads stay balanced
  std::atomic<int64_t> next(0);
  std::atomic<int> bad(0);
  auto work = [&]() {
    for (;;) {
      const int64_t w = next.fetch_add(1);
      if (w >= 3 * count) break;
      const int64_t b = w / 3;
      const int64_t n0 = n_off[b], nn = n_off[b + 1] - n0, e0 = e_off[b], ne = e_off[b + 1] - e0;
      switch (w % 3) {
        case 0: if (eidx32) widen(eidx32 + e0, ne, edge_index + e0); break;   // (NULL: edge_index came as i64)
        case 1: if (!expand_runs(counts + n0, nn, cols + e0, ne)) bad.store(1); break;
        default: widen(samples32 + n0, nn, samples + n0); break;
      }
    }
    fence();
  };
  const int nt = (int)std::min<int64_t>(num_threads, 3 * count);
  std::vector<std::thread> pool;
  pool.reserve((size_t)nt);
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  if (bad.load()) {
    tchgeo::set_last_error("transport: the run lengths of a batch do not add up to its edge count");
    return TCHGEO_ERR_INTERNAL;
  }
  return TCHGEO_OK;
}
