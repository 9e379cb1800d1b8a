// node2vec p/q-biased random walk over CSR for sm_100a.
//
// Replaces src/algo/random_walk.rs:10-75 of the reference (rejection sampling, :52-66, with
// has_edge = binary search in the candidate's sorted adjacency, src/data/graph.rs:80-83).
//
// One thread per walker.  Every attempt draws (neighbour index, f32 uniform) from
// Philox(seed; walker, step, attempt/2), so a walk does not depend on launch geometry or sharding.
// Exact shortcuts that do not change any outcome:
//   * the binary search is skipped when the uniform already decides the attempt
//     (r < min(prob1, prob2) accepts, r >= max(prob1, prob2) rejects, whatever has_edge says);
//   * prev == -1 on the first step never matches (quirk Q8), no search;
//   * the accepted candidate's (row_ptrs[next], row_ptrs[next+1]) pair is carried into the next step.
// Output rows are [walk_length+1] i64 (648 B for length 80): each warp stages 32 walkers x 8 steps
// in shared memory and writes 64-byte runs instead of 8-byte strided stores; dead walkers and the
// -1 padding are written by the same path, so no separate fill(-1) pass over the output is needed.
// Memory behaviour (measured, products-shaped graph): the kernel lives on L1/L2 reuse of a walker's
// current adjacency (the has_edge probes and the next step's gather hit the same lines), so MORE
// resident walkers or a smaller L1 (bigger staging tile) make it slower: 4-5 CTAs/SM and an 8-step
// staging tile are the measured optimum; row_ptrs is kept in L2 (evict_last) and the adjacency reads use
// the optional int32 replica (100 ms -> 63 ms for 24.5 M walkers x 80 steps).
#include <cstdlib>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "graph.cuh"

namespace tchgeo {
namespace {

constexpr int WALK_THREADS = 256;
constexpr uint32_t WALK_MAX_ATTEMPTS = 1u << 25;

struct WalkParams {
  const int64_t* row_ptrs;
  const int64_t* col_indices;
  const int32_t* col_indices32;  // optional int32 replica (half the DRAM lines per adjacency)
  const int64_t* start;
  int64_t* walks;
  unsigned long long* attempts;
  uint32_t* err;
  int64_t num_rows;
  int64_t num_walks;
  int64_t walk_length;
  int64_t walker_base;
  float prob0, prob1, prob2;
  uint32_t key0, key1;
};

template <bool I32>
__device__ __forceinline__ int64_t load_col(const WalkParams& p, int64_t pos) {
  return I32 ? (int64_t)ld_gather64_i32(p.col_indices32 + pos) : ld_gather64_i64(p.col_indices + pos);
}

template <bool I32>
__device__ __forceinline__ bool has_edge_dev(const WalkParams& p, int64_t lo, int64_t hi, int64_t y) {
  while (lo < hi) {  // graph.rs:80-83
    const int64_t mid = lo + ((hi - lo) >> 1);
    const int64_t v = load_col<I32>(p, mid);
    if (v == y) return true;
    if (v < y) lo = mid + 1; else hi = mid;
  }
  return false;
}

template <bool I32, int MINB, int CH>
__global__ void __launch_bounds__(WALK_THREADS, MINB) walk_kernel(const WalkParams p) {
  __shared__ int64_t tile[WALK_THREADS / 32][32][CH + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * WALK_THREADS + threadIdx.x;
  const int64_t warp_first = i - lane;
  const int64_t L = p.walk_length + 1;
  const bool active = i < p.num_walks;
  const uint64_t walker = (uint64_t)(p.walker_base + i);
  const float pmin = fminf(p.prob1, p.prob2), pmax = fmaxf(p.prob1, p.prob2);
  const uint64_t keep = l2_policy_evict_last();  // row_ptrs (8 B/node) should stay in the 126 MB L2

  int64_t cur = active ? p.start[i] : -1;
  int64_t prev = -1;
  int64_t nb = 0, ne = 0, pnb = 0, pne = 0;
  bool alive = active;
  if (alive) {
    if (cur < 0 || cur >= p.num_rows) {
      atomicOr(p.err, DEV_ERR_INDEX);
      alive = false;
    } else {
      nb = ld_gather64_keep_i64(p.row_ptrs + cur, keep);
      ne = ld_gather64_keep_i64(p.row_ptrs + cur + 1, keep);
    }
  }
  unsigned long long my_attempts = 0;

  for (int64_t c0 = 0; c0 < L; c0 += CH) {
    const int nc = (int)min((int64_t)CH, L - c0);
    for (int cc = 0; cc < nc; ++cc) {
      const int64_t colidx = c0 + cc;
      int64_t out;
      if (colidx == 0) {
        out = cur;  // walks[i, 0] = start, random_walk.rs:41
      } else {
        if (alive) {
          const int64_t d = ne - nb;
          if (d <= 0) {
            alive = false;  // random_walk.rs:45-47: the rest of the row stays -1
          } else {
            const uint32_t l = (uint32_t)(colidx - 1);
            int64_t next = -1, nnb = 0, nne = 0;
            Philox4 r4 = {0, 0, 0, 0};
            bool accepted = false;
            for (uint32_t a = 0; a < WALK_MAX_ATTEMPTS; ++a) {
              if ((a & 1u) == 0u)
                r4 = philox4x32_10((uint32_t)walker, (uint32_t)(walker >> 32), l, TAG_WALK | ((a >> 1) << 8), p.key0, p.key1);
              const uint32_t ri = (a & 1u) ? r4.z : r4.x;
              const uint32_t rf = (a & 1u) ? r4.w : r4.y;
              next = load_col<I32>(p, nb + (int64_t)__umulhi(ri, (uint32_t)d));  // :53
              const float r = (float)(rf >> 8) * (1.0f / 16777216.0f);               // :54
              ++my_attempts;
              if (next == prev) {  // :56-58
                if (r < p.prob0) { nnb = pnb; nne = pne; accepted = true; break; }
                continue;
              }
              if (r >= pmax) continue;  // rejected whatever has_edge says
              if (next < 0 || next >= p.num_rows) { atomicOr(p.err, DEV_ERR_INDEX); break; }
              nnb = ld_gather64_keep_i64(p.row_ptrs + next, keep);
              nne = ld_gather64_keep_i64(p.row_ptrs + next + 1, keep);
              if (r < pmin) { accepted = true; break; }  // accepted whatever has_edge says
              const bool he = prev >= 0 && has_edge_dev<I32>(p, nnb, nne, prev);  // :59
              if (he ? (r < p.prob1) : (r < p.prob2)) { accepted = true; break; }        // :60-65
            }
            if (accepted) {
              prev = cur; pnb = nb; pne = ne;  // :68-69
              cur = next; nb = nnb; ne = nne;
            } else {
              if (next >= 0 && next < p.num_rows) atomicOr(p.err, DEV_ERR_WATCHDOG);
              alive = false;
            }
          }
        }
        out = alive ? cur : -1;
      }
      tile[warp][lane][cc] = out;
    }
    __syncwarp();
    // cooperative write-out: 32 rows x nc columns, runs of nc contiguous i64 per row
    for (int idx = lane; idx < 32 * nc; idx += 32) {
      const int r = idx / nc, cc = idx - r * nc;
      const int64_t wi = warp_first + r;
      if (wi < p.num_walks) st_cs_i64(p.walks + wi * L + c0 + cc, tile[warp][r][cc]);
    }
    __syncwarp();
  }
  if (p.attempts) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_attempts += __shfl_xor_sync(0xffffffffu, my_attempts, o);
    if (lane == 0 && my_attempts) atomicAdd(p.attempts, my_attempts);
  }
}


// ---------------------------------------------------------------------------------------------
// Temporal random walk, src/algo/random_walk.rs:80-158.
// One warp per walker.  Every step scans cur's whole adjacency (the reference does the same): an edge's
// timestamp is edge_timestamps[e], or node_timestamps[neighbour] when that is -1 (:118-122); it passes when it
// is -1, when the walk's start timestamp is -1, or when it lies in [start_ts + w0, start_ts + w1) (:125-135).
// The next node is reservoir_sampling with k = 1 over the passing neighbours (:138): the item at passing
// position i >= 1 replaces the pick iff its draw j ~ U[0, i) is 0, so the LAST such item wins, else the first
// passing item -- evaluated in parallel with ballots (passing position) and a warp max.  A step without passing
// neighbours restarts from a uniformly drawn earlier position of the same walk (:140-144).
// Draws: item i uses word i&3 of Philox(walker, step, TAG_TEMPO | (i>>2) << 8); restart uses TAG_TEMPO_RESTART.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t TAG_TEMPO = 6u;
constexpr uint32_t TAG_TEMPO_RESTART = 7u;
constexpr int TEMPO_THREADS = 256;

struct TempoParams {
  const int64_t* row_ptrs;
  const int64_t* col_indices;
  const int64_t* node_ts;
  const int64_t* edge_ts;
  const int64_t* start;
  const int64_t* start_ts;
  int64_t* walks;
  int64_t* walks_ts;
  uint32_t* err;
  int64_t num_rows, num_node_ts, num_walks, L, walker_base, w0, w1;
  uint32_t key0, key1;
};

__global__ void __launch_bounds__(TEMPO_THREADS) tempo_walk_kernel(const TempoParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * TEMPO_THREADS + threadIdx.x) >> 5;
  if (i >= p.num_walks) return;
  const uint64_t walker = (uint64_t)(p.walker_base + i);
  const uint32_t wlo = (uint32_t)walker, whi = (uint32_t)(walker >> 32);
  int64_t* row = p.walks + i * p.L;
  int64_t* row_ts = p.walks_ts + i * p.L;
  int64_t cur = p.start[i];
  const int64_t i_ts = p.start_ts[i];
  const int64_t lo_w = i_ts + p.w0, hi_w = i_ts + p.w1;  // Range: lo_w <= t < hi_w
  if (lane == 0) { row[0] = cur; row_ts[0] = i_ts; }
  bool dead = false;
  for (int64_t l = 0; l + 1 < p.L; ++l) {
    int64_t next = -1, next_ts = -1;
    if (!dead) {
      if (cur < 0 || cur >= p.num_rows) {  // neighbors_range(cur) out of bounds: the reference panics
        if (lane == 0) atomicOr(p.err, DEV_ERR_INDEX);
        dead = true;
      }
    }
    if (!dead) {
      const int64_t b = __ldg(p.row_ptrs + cur), e = __ldg(p.row_ptrs + cur + 1);
      uint32_t npass = 0;                 // passing neighbours seen so far (warp-uniform)
      uint32_t best = 0;                  // largest passing position >= 1 whose draw hit, 0 = none (per lane)
      bool has_first = false;             // this lane holds passing position 0
      int64_t best_node = -1, best_ts = -1, first_node = -1, first_ts = -1;
      for (int64_t base = b; base < e; base += 32) {
        const int64_t ep = base + lane;
        bool pass = false;
        int64_t node = -1, ts = -1;
        if (ep < e) {
          node = __ldg(p.col_indices + ep);
          ts = __ldg(p.edge_ts + ep);
          if (ts == -1) {  // NAN_TIMESTAMP: fall back to the neighbour's own timestamp, :118-122
            if (node < 0 || node >= p.num_node_ts) atomicOr(p.err, DEV_ERR_INDEX);
            else ts = __ldg(p.node_ts + node);
          }
          pass = ts == -1 || i_ts == -1 || (lo_w <= ts && ts < hi_w);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (pass) {
          const uint32_t pos = npass + (uint32_t)__popc(m & ((1u << lane) - 1u));
          if (pos == 0) {
            has_first = true; first_node = node; first_ts = ts;
          } else {
            const Philox4 r = philox4x32_10(wlo, whi, (uint32_t)l, TAG_TEMPO | ((pos >> 2) << 8), p.key0, p.key1);
            if (__umulhi(pick4(r, pos & 3u), pos) == 0u) { best = pos; best_node = node; best_ts = ts; }  // later wins
          }
        }
        npass += (uint32_t)__popc(m);
      }
      if (npass >= (1u << 24)) { if (lane == 0) atomicOr(p.err, DEV_ERR_INDEX); dead = true; }
      if (!dead && npass > 0) {
        uint32_t top = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        // the winner is the passing item at position `top`; top == 0: no later item hit, the first one stays
        const uint32_t om = __ballot_sync(0xffffffffu, top ? (best == top) : has_first);
        const int src = __ffs(om) - 1;
        next = __shfl_sync(0xffffffffu, top ? best_node : first_node, src);
        next_ts = __shfl_sync(0xffffffffu, top ? best_ts : first_ts, src);
      } else if (!dead) {
        // restart: a uniformly drawn earlier position of this walk (lane 0 re-reads what it wrote), :140-144
        const Philox4 r = philox4x32_10(wlo, whi, (uint32_t)l, TAG_TEMPO_RESTART, p.key0, p.key1);
        const int64_t ri = (int64_t)__umulhi(r.x, (uint32_t)(l + 1));
        if (lane == 0) { next = row[ri]; next_ts = row_ts[ri]; }
        next = __shfl_sync(0xffffffffu, next, 0);
        next_ts = __shfl_sync(0xffffffffu, next_ts, 0);
      }
    }
    if (dead) { next = -1; next_ts = -1; }
    cur = next;
    if (lane == 0) { row[l + 1] = cur; row_ts[l + 1] = next_ts; }
  }
}


// ---- thread-per-walker form (the default) ------------------------------------------------------------------------------
// The warp-per-walker kernel above issues ~200 warp instructions per step for a 25-neighbour node (one Philox evaluation
// per 32-neighbour chunk, ballots, a warp max): 12.7 G warp instructions, 59 % issue utilisation, 1.3 TB/s of DRAM for the
// products-shaped graph -- instruction-bound.  Here every THREAD owns a walker and walks its current node's adjacency
// serially: the reservoir becomes "overwrite on a hit" (the last hit wins by construction), one Philox evaluation serves
// four consecutive passing positions, and an instruction serves 32 walkers.  A walker standing on a node with more than
// TEMPO_LIGHT neighbours is handed to the whole warp (the chunked scan of the kernel above), one such walker at a time.
// Same draws, same results, bit for bit.
constexpr int TEMPO_LIGHT = 96;

struct TempoPick {
  int64_t node, ts;
  uint32_t npass;
};

// the warp scans [b, e) for the walker of lane `src` (all its parameters are broadcast from that lane)
__device__ __noinline__ TempoPick tempo_scan_warp(const TempoParams& p, int lane, int64_t b, int64_t e, int64_t i_ts,
                                                     int64_t lo_w, int64_t hi_w, uint32_t wlo, uint32_t whi, uint32_t l,
                                                     bool& bad) {
  uint32_t npass = 0, best = 0;
  bool has_first = false;
  int64_t best_node = -1, best_ts = -1, first_node = -1, first_ts = -1;
  for (int64_t base = b; base < e; base += 32) {
    const int64_t ep = base + lane;
    bool pass = false;
    int64_t node = -1, ts = -1;
    if (ep < e) {
      node = __ldg(p.col_indices + ep);
      ts = __ldg(p.edge_ts + ep);
      if (ts == -1) {
        if (node < 0 || node >= p.num_node_ts) bad = true;
        else ts = __ldg(p.node_ts + node);
      }
      pass = ts == -1 || i_ts == -1 || (lo_w <= ts && ts < hi_w);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (pass) {
      const uint32_t pos = npass + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (pos == 0) {
        has_first = true; first_node = node; first_ts = ts;
      } else {
        const Philox4 r = philox4x32_10(wlo, whi, l, TAG_TEMPO | ((pos >> 2) << 8), p.key0, p.key1);
        if (__umulhi(pick4(r, pos & 3u), pos) == 0u) { best = pos; best_node = node; best_ts = ts; }
      }
    }
    npass += (uint32_t)__popc(m);
  }
  TempoPick out{-1, -1, npass};
  if (npass > 0) {
    uint32_t top = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
    const uint32_t om = __ballot_sync(0xffffffffu, top ? (best == top) : has_first);
    const int from = __ffs(om) - 1;
    out.node = __shfl_sync(0xffffffffu, top ? best_node : first_node, from);
    out.ts = __shfl_sync(0xffffffffu, top ? best_ts : first_ts, from);
  }
  return out;
}

__global__ void __launch_bounds__(TEMPO_THREADS, 4) tempo_walk_thread_kernel(const TempoParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * TEMPO_THREADS + threadIdx.x;
  const bool mine = i < p.num_walks;            // (idle lanes of the last warp still help with the hubs)
  const uint64_t walker = (uint64_t)(p.walker_base + (mine ? i : 0));
  const uint32_t wlo = (uint32_t)walker, whi = (uint32_t)(walker >> 32);
  int64_t* row = p.walks + (mine ? i : 0) * p.L;
  int64_t* row_ts = p.walks_ts + (mine ? i : 0) * p.L;
  int64_t cur = mine ? p.start[i] : -1;
  const int64_t i_ts = mine ? p.start_ts[i] : 0;
  const int64_t lo_w = i_ts + p.w0, hi_w = i_ts + p.w1;  // Range: lo_w <= t < hi_w
  if (mine) { row[0] = cur; row_ts[0] = i_ts; }
  bool dead = !mine, bad = false;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.col_indices) | reinterpret_cast<uintptr_t>(p.edge_ts)) & 31u) == 0;
  for (int64_t l = 0; l + 1 < p.L; ++l) {
    int64_t next = -1, next_ts = -1, b = 0, e = 0;
    uint32_t npass = 0;
    if (!dead && (cur < 0 || cur >= p.num_rows)) {  // neighbors_range(cur) out of bounds: the reference panics
      bad = true;
      dead = true;
    }
    if (!dead) {
      b = __ldg(p.row_ptrs + cur);
      e = __ldg(p.row_ptrs + cur + 1);
    }
    const bool heavy = !dead && e - b > TEMPO_LIGHT;
    if (!dead && !heavy) {
      Philox4 r{0u, 0u, 0u, 0u};
      uint32_t rblk = 0xFFFFFFFFu;
      // The adjacency is read in ALIGNED groups of four neighbours (one 32-byte sector of ids, one of timestamps), all
      // of a group's loads issued before its first neighbour is looked at: every sector is fetched by one pair of
      // 16-byte loads instead of surviving four loop iterations in an L1 that a thousand such streams overflow.
      for (int64_t g = b & ~(int64_t)3; g < e; g += 4) {
        int64_t nd[4], tt[4];
        if (vec_ok && g >= b && g + 4 <= e) {
          const longlong2 c0 = __ldg(reinterpret_cast<const longlong2*>(p.col_indices + g));
          const longlong2 c1 = __ldg(reinterpret_cast<const longlong2*>(p.col_indices + g) + 1);
          const longlong2 t0 = __ldg(reinterpret_cast<const longlong2*>(p.edge_ts + g));
          const longlong2 t1 = __ldg(reinterpret_cast<const longlong2*>(p.edge_ts + g) + 1);
          nd[0] = c0.x; nd[1] = c0.y; nd[2] = c1.x; nd[3] = c1.y;
          tt[0] = t0.x; tt[1] = t0.y; tt[2] = t1.x; tt[3] = t1.y;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool in = g + u >= b && g + u < e;
            nd[u] = in ? __ldg(p.col_indices + g + u) : -1;
            tt[u] = in ? __ldg(p.edge_ts + g + u) : 0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (g + u < b || g + u >= e) continue;
          const int64_t node = nd[u];
          int64_t ts = tt[u];
          if (ts == -1) {  // NAN_TIMESTAMP: fall back to the neighbour's own timestamp, :118-122
            if (node < 0 || node >= p.num_node_ts) bad = true;
            else ts = __ldg(p.node_ts + node);
          }
          if (!(ts == -1 || i_ts == -1 || (lo_w <= ts && ts < hi_w))) continue;
          const uint32_t pos = npass++;
          if (pos == 0) {
            next = node; next_ts = ts;
          } else {
            if ((pos >> 2) != rblk) {
              rblk = pos >> 2;
              r = philox4x32_10(wlo, whi, (uint32_t)l, TAG_TEMPO | (rblk << 8), p.key0, p.key1);
            }
            if (__umulhi(pick4(r, pos & 3u), pos) == 0u) { next = node; next_ts = ts; }   // a later hit replaces the pick
          }
        }
      }
    }
    // hubs: the warp scans them together, one walker at a time
    uint32_t hm = __ballot_sync(0xffffffffu, heavy);
    while (hm) {
      const int src = __ffs(hm) - 1;
      hm &= hm - 1;
      bool wbad = false;
      const TempoPick pk = tempo_scan_warp(p, lane, __shfl_sync(0xffffffffu, b, src), __shfl_sync(0xffffffffu, e, src),
                                           __shfl_sync(0xffffffffu, i_ts, src), __shfl_sync(0xffffffffu, lo_w, src),
                                           __shfl_sync(0xffffffffu, hi_w, src), __shfl_sync(0xffffffffu, wlo, src),
                                           __shfl_sync(0xffffffffu, whi, src), (uint32_t)l, wbad);
      bad |= wbad;
      if (lane == src) { next = pk.node; next_ts = pk.ts; npass = pk.npass; }
    }
    if (!dead) {
      if (npass >= (1u << 24)) {
        bad = true;
        dead = true;
      } else if (npass == 0) {
        // restart: a uniformly drawn earlier position of this walk (the thread re-reads what it wrote), :140-144
        const Philox4 r = philox4x32_10(wlo, whi, (uint32_t)l, TAG_TEMPO_RESTART, p.key0, p.key1);
        const int64_t ri = (int64_t)__umulhi(r.x, (uint32_t)(l + 1));
        next = row[ri]; next_ts = row_ts[ri];
      }
    }
    if (dead) { next = -1; next_ts = -1; }
    cur = next;
    if (mine) { row[l + 1] = cur; row_ts[l + 1] = next_ts; }
  }
  if (bad) atomicOr(p.err, DEV_ERR_INDEX);
}

}  // namespace
}  // namespace tchgeo

using namespace tchgeo;

extern "C" tchgeo_status tchgeo_random_walk(const int64_t* row_ptrs, int64_t num_rows, const int64_t* col_indices,
                                            const int64_t* start, int64_t num_walks, int64_t walk_length, float p,
                                            float q, uint64_t seed, int64_t walker_base, int64_t* walks,
                                            int64_t* stats, int64_t* attempts_out, tchgeo_stream stream_) {
  return tchgeo_random_walk_ex(row_ptrs, num_rows, col_indices, nullptr, start, num_walks, walk_length, p, q, seed,
                               walker_base, walks, stats, attempts_out, stream_);
}

extern "C" tchgeo_status tchgeo_random_walk_graph(const tchgeo_graph_t* graph, int32_t rel, const int64_t* start,
                                                  int64_t num_walks, int64_t walk_length, float p, float q, uint64_t seed,
                                                  int64_t walker_base, int64_t* walks, int64_t* stats,
                                                  int64_t* attempts_out, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(graph != nullptr && rel >= 0 && rel < graph->R, "bad graph handle / relation");
  const tchgeo_status st = graph_ensure(const_cast<tchgeo_graph*>(graph), TCHGEO_PREPARE_INDEX_REPLICA, (cudaStream_t)stream_);
  if (st != TCHGEO_OK) return st;
  const size_t r = (size_t)rel;
  return tchgeo_random_walk_ex(graph->ptrs[r], graph->num_major[r], graph->indices[r], graph->indices32[r], start, num_walks,
                               walk_length, p, q, seed, walker_base, walks, stats, attempts_out, stream_);
}

extern "C" tchgeo_status tchgeo_random_walk_ex(const int64_t* row_ptrs, int64_t num_rows, const int64_t* col_indices,
                                               const int32_t* col_indices32, const int64_t* start, int64_t num_walks,
                                               int64_t walk_length, float p, float q, uint64_t seed,
                                               int64_t walker_base, int64_t* walks, int64_t* stats,
                                               int64_t* attempts_out, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_walks >= 0 && walk_length >= 0 && num_rows >= 0, "negative size");
  TCHGEO_REQUIRE(row_ptrs && stats && (num_walks == 0 || (start && walks)), "NULL pointer");
  TCHGEO_REQUIRE(walk_length < ((int64_t)1 << 31), "walk_length too large");
  // "p or q may not be 0 or nan", random_walk.rs:30
  TCHGEO_REQUIRE(p == p && q == q && p != 0.0f && q != 0.0f, "p or q may not be 0 or nan");
  if (attempts_out) *attempts_out = 0;
  if (num_walks == 0) return TCHGEO_OK;
  // random_walk.rs:29-36, f32 arithmetic; max_by keeps the last maximum
  const float a = 1.0f / p, b = 1.0f, c = 1.0f / q;
  float max_prob = a;
  if (b >= max_prob) max_prob = b;
  if (c >= max_prob) max_prob = c;
  WalkParams wp;
  wp.row_ptrs = row_ptrs; wp.col_indices = col_indices; wp.col_indices32 = col_indices32; wp.start = start;
  wp.walks = walks;
  wp.attempts = (unsigned long long*)stats;
  wp.err = (uint32_t*)(stats + 1);
  wp.num_rows = num_rows; wp.num_walks = num_walks; wp.walk_length = walk_length; wp.walker_base = walker_base;
  wp.prob0 = 1.0f / p / max_prob;
  wp.prob1 = 1.0f / max_prob;
  wp.prob2 = 1.0f / q / max_prob;
  wp.key0 = (uint32_t)seed; wp.key1 = (uint32_t)(seed >> 32);
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(stats, 0, 16, stream));
  const int64_t grid = (num_walks + WALK_THREADS - 1) / WALK_THREADS;
  TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many walkers for one launch");
  static int minb = -1;  // tuning knob TCHGEO_WALK_MIN_BLOCKS = 4 | 5 | 6 (register budget 56 / 48 / 40)
  if (minb < 0) {
    const char* e = getenv("TCHGEO_WALK_MIN_BLOCKS");
    const int x = e ? atoi(e) : 5;
    minb = (x >= 4 && x <= 6) ? x : 5;
  }
  static int chunk = -1;  // tuning knob TCHGEO_WALK_CHUNK = 4 | 8 | 16 staged steps (shared memory vs L1 size)
  if (chunk < 0) {
    const char* e = getenv("TCHGEO_WALK_CHUNK");
    const int x = e ? atoi(e) : 8;
    chunk = (x == 4 || x == 8 || x == 16) ? x : 8;
  }
#define TCHGEO_LAUNCH_WALK(I32, MB)                                                                  \
  do {                                                                                               \
    if (chunk == 4) walk_kernel<I32, MB, 4><<<(unsigned)grid, WALK_THREADS, 0, stream>>>(wp);        \
    else if (chunk == 16) walk_kernel<I32, MB, 16><<<(unsigned)grid, WALK_THREADS, 0, stream>>>(wp); \
    else walk_kernel<I32, MB, 8><<<(unsigned)grid, WALK_THREADS, 0, stream>>>(wp);                   \
  } while (0)
  if (col_indices32) {
    if (minb == 4) TCHGEO_LAUNCH_WALK(true, 4); else if (minb == 6) TCHGEO_LAUNCH_WALK(true, 6); else TCHGEO_LAUNCH_WALK(true, 5);
  } else {
    if (minb == 4) TCHGEO_LAUNCH_WALK(false, 4); else if (minb == 6) TCHGEO_LAUNCH_WALK(false, 6); else TCHGEO_LAUNCH_WALK(false, 5);
  }
#undef TCHGEO_LAUNCH_WALK
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  int64_t h[2] = {0, 0};
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(h, stats, 16, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (attempts_out) *attempts_out = h[0];
  return status_from_dev_err((uint32_t)h[1]);
}

extern "C" tchgeo_status tchgeo_tempo_random_walk(const int64_t* row_ptrs, int64_t num_rows, const int64_t* col_indices,
                                                  const int64_t* node_timestamps, int64_t num_node_timestamps,
                                                  const int64_t* edge_timestamps, const int64_t* start,
                                                  const int64_t* start_timestamps, int64_t num_walks, int64_t walk_length,
                                                  int64_t window_lo, int64_t window_hi, uint64_t seed, int64_t walker_base,
                                                  int64_t* walks, int64_t* walks_timestamps, int32_t* scratch,
                                                  tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_walks >= 0 && num_rows >= 0 && num_node_timestamps >= 0, "negative size");
  TCHGEO_REQUIRE(walk_length >= 0 && walk_length < ((int64_t)1 << 31), "walk_length out of range");
  TCHGEO_REQUIRE(scratch != nullptr, "NULL pointer");
  if (num_walks == 0) return TCHGEO_OK;
  if (walk_length == 0) {  // walks_data[i * L] on an empty tensor: index panic, random_walk.rs:113
    set_last_error("walk_length 0 with a non-empty start (the reference panics)");
    return TCHGEO_ERR_REFERENCE_PANIC;
  }
  TCHGEO_REQUIRE(row_ptrs && start && start_timestamps && walks && walks_timestamps, "NULL pointer");
  TempoParams tp;
  tp.row_ptrs = row_ptrs; tp.col_indices = col_indices; tp.node_ts = node_timestamps; tp.edge_ts = edge_timestamps;
  tp.start = start; tp.start_ts = start_timestamps; tp.walks = walks; tp.walks_ts = walks_timestamps;
  tp.err = (uint32_t*)scratch;
  tp.num_rows = num_rows; tp.num_node_ts = num_node_timestamps; tp.num_walks = num_walks; tp.L = walk_length;
  tp.walker_base = walker_base; tp.w0 = window_lo; tp.w1 = window_hi;
  tp.key0 = (uint32_t)seed; tp.key1 = (uint32_t)(seed >> 32);
  cudaStream_t stream = (cudaStream_t)stream_;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4, stream));
  const char* form = getenv("TCHGEO_TEMPO_WALK");   // "warp": the warp-per-walker kernel (kept for comparison)
  if (form && strcmp(form, "warp") == 0) {
    const int64_t grid = (num_walks * 32 + TEMPO_THREADS - 1) / TEMPO_THREADS;
    TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many walkers for one launch");
    tempo_walk_kernel<<<(unsigned)grid, TEMPO_THREADS, 0, stream>>>(tp);
  } else {
    const int64_t grid = (num_walks + TEMPO_THREADS - 1) / TEMPO_THREADS;
    TCHGEO_REQUIRE(grid < ((int64_t)1 << 31), "too many walkers for one launch");
    tempo_walk_thread_kernel<<<(unsigned)grid, TEMPO_THREADS, 0, stream>>>(tp);
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  uint32_t h = 0;
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(&h, scratch, 4, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  return status_from_dev_err(h);
}
