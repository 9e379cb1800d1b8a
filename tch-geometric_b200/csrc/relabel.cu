// Frontier dedup + insertion-order local-id relabeling of sampled trees (additive stage, K7), batched.
//
// Semantic of src/algo/negative_sampling.rs:20-47 (samples_mapping), per tree:
//   nodes     = seeds (all, duplicates kept, :25) ++ every non-seed id at its first appearance (:36-39)
//   map[seed] = index of the seed's LAST occurrence (HashMap::extend overwrites, :26)
//   local[i]  = map[samples[i]]
//
// Parallel formulation.  Every position i of a tree gets a priority
//     prio(i) = i < S ? S-1-i : i          (S = number of seeds)
// so that the minimum priority over the occurrences of an id is exactly the occurrence the serial HashMap ends up
// pointing at: any seed beats any non-seed, the LAST seed beats the earlier ones, the FIRST non-seed beats the later
// ones.  One open-addressing insert (atomicCAS on the key, atomicMin on the priority) resolves all ids of a tree;
// "emits a node" flags (every seed; a non-seed iff it holds its id's minimum) are compacted by an exclusive scan in
// position order, which reproduces the serial insertion order; winners publish their rank and a lookup gives local[].
//
// HBM layout: the trees of a step (B x ~0.6 M ids for the products configuration) do not fit a cache, but the hash
// table of ONE tree does (2^20 slots x 8 B).  The batches are therefore processed in WAVES of as many trees as keep
// tables + ids inside the 126 MB L2 (TCHGEO_RELABEL_WAVE_MB, default 64): per wave one 0xFF memset of the tables and
// three kernels (insert, flag + scan + compact, lookup), whose table accesses and re-reads of the ids are L2 hits; DRAM
// sees 8 B read + (8 B local + <= 8 B nodes) written per id.  Integer work, HBM/L2-bound: no tensor cores.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace tchgeo {
namespace {

inline size_t rl_align(size_t x) { return (x + 255) / 256 * 256; }

constexpr int RL_THREADS = 256;
constexpr int RL_ITEMS = 4;                        // positions per thread in the compaction kernel
constexpr int RL_TILE = RL_THREADS * RL_ITEMS;     // positions per tile of the per-tree scan
constexpr uint32_t RL_NOSLOT = 0xFFFFFFFFu;
constexpr uint64_t RL_ST_AGG = 1ull << 62;         // look-back status flags; 3 (= the 0xFF fill) means "not published"
constexpr uint64_t RL_ST_INCL = 2ull << 62;
constexpr uint64_t RL_ST_MASK = (1ull << 62) - 1;

// One hash table per tree of the wave.  K32: ids < 2^32-1, slot = (key, prio) in one 8-byte word pair, so the CAS and the
// min of an insert touch one sector.  K64: any non-negative i64 id, keys and priorities in two arrays.
template <bool K32>
struct Table;
template <>
struct Table<true> {
  uint2* ent;
  static constexpr size_t slot_bytes = 8;
  __device__ __forceinline__ static bool fits(int64_t key) { return (uint64_t)key < 0xFFFFFFFFull; }
  __device__ __forceinline__ uint32_t home(int64_t key, uint32_t mask, int shift) const {
    return (((uint32_t)key * 0x9E3779B1u) >> shift) & mask;
  }
  // claims or finds the slot of `key`, starting at h
  __device__ __forceinline__ uint32_t insert(int64_t key, uint32_t h, uint32_t mask) const {
    const uint32_t k = (uint32_t)key;
    while (true) {
      uint32_t cur = ent[h].x;
      if (cur == 0xFFFFFFFFu) cur = atomicCAS(&ent[h].x, 0xFFFFFFFFu, k);
      if (cur == 0xFFFFFFFFu || cur == k) return h;
      h = (h + 1) & mask;
    }
  }
  __device__ __forceinline__ void min_prio(uint32_t h, uint32_t prio) const { atomicMin(&ent[h].y, prio); }
  __device__ __forceinline__ uint32_t prio(uint32_t h) const { return ent[h].y; }
};
template <>
struct Table<false> {
  unsigned long long* keys;
  uint32_t* prios;
  static constexpr size_t slot_bytes = 12;
  __device__ __forceinline__ static bool fits(int64_t key) { return key >= 0; }
  __device__ __forceinline__ uint32_t home(int64_t key, uint32_t mask, int shift) const {
    const uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(h >> (32 + shift)) & mask;
  }
  __device__ __forceinline__ uint32_t insert(int64_t key, uint32_t h, uint32_t mask) const {
    const unsigned long long k = (unsigned long long)key;
    while (true) {
      unsigned long long cur = keys[h];
      if (cur == ~0ull) cur = atomicCAS(keys + h, ~0ull, k);
      if (cur == ~0ull || cur == k) return h;
      h = (h + 1) & mask;
    }
  }
  __device__ __forceinline__ void min_prio(uint32_t h, uint32_t prio) const { atomicMin(prios + h, prio); }
  __device__ __forceinline__ uint32_t prio(uint32_t h) const { return prios[h]; }
};

struct RlParams {
  const int64_t* samples;   // [B, stride]
  int64_t stride;
  const int64_t* lens;      // DEVICE [B] ids per tree (clamped to n_max)
  int64_t* nodes;           // [B, stride]
  int64_t* local;           // [B, stride]
  int64_t* nodes_len;       // DEVICE [B]
  int64_t num_seeds;
  int64_t n_max;            // bound of lens: geometry of slot_of and of the grid
  int32_t b0;               // first tree of the wave
  int32_t tiles_per_tree;   // ceil(n_max / RL_TILE)
  uint32_t cap_mask;        // table slots - 1 (power of two, > n_max)
  int32_t hash_shift;       // 32 - log2(slots)
  char* tables;             // wave slot y: tables + y * table_bytes
  size_t table_bytes;
  uint32_t* rank;           // [wave, slots] rank of the winner of every id
  uint32_t* slot_of;        // [wave, n_max] table slot of every position
  uint64_t* status;         // [wave, tiles_per_tree] look-back words (0xFF-filled = not published)
  uint32_t* ticket;         // tile dispenser of this wave's compaction kernel (zero)
  uint32_t* err;
};

template <bool K32>
__device__ __forceinline__ Table<K32> table_of(const RlParams& p, int y);
template <>
__device__ __forceinline__ Table<true> table_of<true>(const RlParams& p, int y) {
  return Table<true>{reinterpret_cast<uint2*>(p.tables + (size_t)y * p.table_bytes)};
}
template <>
__device__ __forceinline__ Table<false> table_of<false>(const RlParams& p, int y) {
  char* base = p.tables + (size_t)y * p.table_bytes;
  return Table<false>{reinterpret_cast<unsigned long long*>(base),
                      reinterpret_cast<uint32_t*>(base + ((size_t)p.cap_mask + 1) * 8)};
}

__device__ __forceinline__ int64_t rl_len(const RlParams& p, int b) {
  int64_t n = p.lens[b];
  if (n > p.n_max) n = p.n_max;
  return n < 0 ? 0 : n;
}

// ---- pass 1: insert every id of the wave's trees, keep the minimum priority per id -----------------------
template <bool K32>
__global__ void __launch_bounds__(RL_THREADS) rl_insert_kernel(const RlParams p) {
  const int y = blockIdx.y, b = p.b0 + y;
  const int64_t n = rl_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * RL_TILE;
  if (i0 >= n) return;
  const Table<K32> tab = table_of<K32>(p, y);
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  uint32_t* slot_of = p.slot_of + (int64_t)y * p.n_max;
  int64_t key[RL_ITEMS];
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {  // coalesced: consecutive threads read consecutive ids
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    key[u] = i < n ? __ldg(src + i) : -1;
  }
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    if (i >= n) continue;
    if (!Table<K32>::fits(key[u])) {
      atomicOr(p.err, DEV_ERR_INDEX);
      slot_of[i] = RL_NOSLOT;
      continue;
    }
    const uint32_t h = tab.insert(key[u], tab.home(key[u], p.cap_mask, p.hash_shift), p.cap_mask);
    tab.min_prio(h, i < p.num_seeds ? (uint32_t)(p.num_seeds - 1 - i) : (uint32_t)i);
    slot_of[i] = h;
  }
}

// ---- pass 2: flags, exclusive scan in position order (decoupled look-back per tree), compaction, ranks ----
template <bool K32>
__global__ void __launch_bounds__(RL_THREADS) rl_compact_kernel(const RlParams p) {
  __shared__ uint32_t s_wtot[RL_THREADS / 32];
  __shared__ uint32_t s_tile;
  __shared__ int64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles are dispensed in start order, tree-major: every tile a look-back waits for (same tree, smaller index)
  // holds a smaller ticket, i.e. belongs to a CTA that is already running
  if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const int y = (int)(s_tile / (uint32_t)p.tiles_per_tree), t = (int)(s_tile - (uint32_t)y * (uint32_t)p.tiles_per_tree);
  const int b = p.b0 + y;
  const int64_t n = rl_len(p, b);
  const int64_t i0 = (int64_t)t * RL_TILE;
  if (i0 >= n && t > 0) return;  // tiles past the end publish nothing: nobody looks back at them
  const Table<K32> tab = table_of<K32>(p, y);
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  const uint32_t* slot_of = p.slot_of + (int64_t)y * p.n_max;
  uint32_t* rank_tab = p.rank + (size_t)y * ((size_t)p.cap_mask + 1);
  const int64_t S = p.num_seeds;

  // blocked arrangement: thread `tid` owns positions i0 + 4*tid .. +3, so ranks follow from one scan of thread sums
  const int64_t ibase = i0 + (int64_t)tid * RL_ITEMS;
  uint32_t slot[RL_ITEMS], pr[RL_ITEMS];
  uint32_t flags = 0, cnt = 0;
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    const int64_t i = ibase + u;
    slot[u] = i < n ? slot_of[i] : RL_NOSLOT;
  }
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) pr[u] = slot[u] != RL_NOSLOT ? tab.prio(slot[u]) : 0u;
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    const int64_t i = ibase + u;
    // every seed is kept (:25); a non-seed emits a node iff it is the first occurrence of an id no seed carries (:36-39)
    const bool f = i < n && (i < S || (slot[u] != RL_NOSLOT && pr[u] == (uint32_t)i));
    flags |= (uint32_t)f << u;
    cnt += f;
  }
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t excl = incl - cnt, total = 0;
#pragma unroll
  for (int w = 0; w < RL_THREADS / 32; ++w) {
    const uint32_t v = s_wtot[w];
    if (w < warp) excl += v;
    total += v;
  }
  uint64_t* st = p.status + (size_t)y * p.tiles_per_tree;
  if (tid == 0) st_relaxed_u64(st + t, (t == 0 ? RL_ST_INCL : RL_ST_AGG) | (uint64_t)total);
  if (warp == 0) {
    int64_t before = 0;
    if (t > 0) {
      int j = t - 1;
      uint32_t spins = 0;
      while (true) {
        const int idx = j - lane;
        const uint64_t v = idx >= 0 ? ld_relaxed_u64(st + idx) : RL_ST_INCL;
        const uint32_t flag = (uint32_t)(v >> 62);
        const uint32_t incl_mask = __ballot_sync(0xffffffffu, flag == 2u);
        const uint32_t inval_mask = __ballot_sync(0xffffffffu, flag == 3u);
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
        int64_t val = lane <= first_incl ? (int64_t)(v & RL_ST_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        before += val;
        if (first_incl < 32) break;
        j -= 32;
      }
      if (lane == 0) st_relaxed_u64(st + t, RL_ST_INCL | (uint64_t)(before + total));
    }
    if (lane == 0) s_excl = before;
  }
  __syncthreads();
  const int64_t tile_excl = s_excl;
  if (tid == 0 && i0 + RL_TILE >= n) p.nodes_len[b] = tile_excl + total;  // the tree's last tile (or its empty tile 0)
  int64_t* nodes = p.nodes + (int64_t)b * p.stride;
  uint32_t r = (uint32_t)tile_excl + excl;
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    if (!((flags >> u) & 1u)) continue;
    const int64_t i = ibase + u;
    st_cs_i64(nodes + r, __ldg(src + i));
    // the occurrence the map points at publishes its rank: the last seed of an id, else its first non-seed
    if (slot[u] != RL_NOSLOT && pr[u] == (i < S ? (uint32_t)(S - 1 - i) : (uint32_t)i)) rank_tab[slot[u]] = r;
    ++r;
  }
}

// ---- pass 3: local[i] = rank of the id's winner ---------------------------------------------------------------
__global__ void __launch_bounds__(RL_THREADS) rl_lookup_kernel(const RlParams p) {
  const int y = blockIdx.y, b = p.b0 + y;
  const int64_t n = rl_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * RL_TILE;
  if (i0 >= n) return;
  const uint32_t* slot_of = p.slot_of + (int64_t)y * p.n_max;
  const uint32_t* rank_tab = p.rank + (size_t)y * ((size_t)p.cap_mask + 1);
  int64_t* local = p.local + (int64_t)b * p.stride;
  uint32_t h[RL_ITEMS];
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    h[u] = i < n ? slot_of[i] : RL_NOSLOT;
  }
#pragma unroll
  for (int u = 0; u < RL_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    if (i < n) st_cs_i64(local + i, h[u] != RL_NOSLOT ? (int64_t)rank_tab[h[u]] : -1);
  }
}

// =================================================================================================
// Persistent form (the one the sampling plan uses when the ids fit 32 bits): ONE cooperative launch per call.
// The grid is G groups of C co-resident CTAs; group g owns one hash table and walks trees g, g+G, ... through
//   clear -> insert -> count -> assign -> lookup
// with a group barrier between the phases (an arrive counter + generation word per group: only the C CTAs that share
// a tree ever wait for each other).  At any moment only G tables (8 MB each for the products configuration), G slot
// maps and the G trees in flight are live, so every table access and every re-read of the ids is an L2 hit, and DRAM
// sees 8 B read + (8 B local + <= 8 B nodes) written per id.  One table slot is ONE 64-bit word (key << 32 | priority),
// so the common insert (a new id) is a single atomicCAS; a second occurrence adds one 64-bit atomicMin (equal keys:
// the smaller priority wins); the winner later overwrites the priority with (1 << 31 | rank) in place.
// Everything the CTAs of a group hand to each other (table, slot map, state bytes, CTA counts) is read with ld.cg:
// the L1 is not coherent and the same addresses are reused tree after tree.
// The co-residency the barriers rely on is what cudaLaunchCooperativeKernel guarantees (two plans on two streams would
// otherwise be able to starve each other's CTAs).
// =================================================================================================
constexpr int RP_THREADS = 512;
constexpr int RP_ITEMS = 4;                       // consecutive positions per thread in the scan phases
constexpr int RP_TILE = RP_THREADS * RP_ITEMS;
constexpr int RP_MAX_GROUPS = 32;
constexpr int RP_MAX_CTAS = 4096;                 // bound of the grid (cta_count scratch)
constexpr unsigned long long RP_EMPTY = ~0ull;
constexpr uint32_t RP_RANKED = 0x80000000u;
// per-position state byte written by the count phase
constexpr uint32_t RP_F_NODE = 1u;      // emits a node (every seed; a non-seed holding its id's minimum)
constexpr uint32_t RP_F_PUBLISH = 2u;   // first occurrence of a non-seed id: publishes its rank in the table
constexpr uint32_t RP_F_PENDING = 4u;   // later occurrence of a non-seed id: local[] comes from the published rank

struct RpParams {
  const int64_t* samples;
  int64_t stride;
  const int64_t* lens;
  int64_t* nodes;
  int64_t* local;
  int64_t* nodes_len;
  int64_t num_seeds, n_max, n_pad;   // n_pad: n_max rounded up to a multiple of RP_ITEMS (row pitch of the scratch maps)
  int32_t num_trees, groups, ctas_per_group;
  uint32_t cap_mask;
  int32_t hash_shift;
  unsigned long long* tables;        // [groups, slots]
  uint32_t* slot_of;                 // [groups, n_pad]
  uint8_t* fbytes;                   // [groups, n_pad]
  uint32_t* cta_count;               // [groups, ctas_per_group]
  uint32_t* bars;                    // [groups, 32]: [0] arrivals, [1] generation (zero-initialised)
  uint32_t* err;
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// barrier over the `n` CTAs of one group (all co-resident).  Returns false when the watchdog trips.
__device__ __forceinline__ bool group_sync(uint32_t* bar, uint32_t n, uint32_t* err) {
  __shared__ uint32_t s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ok = 1u;
    const uint32_t gen = ld_acquire_u32(bar + 1);
    __threadfence();
    if (atomicAdd(bar, 1u) == n - 1u) {
      atomicExch(bar, 0u);
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      uint32_t spins = 0;
      while (ld_acquire_u32(bar + 1) == gen) {
        if (++spins > (1u << 26)) {
          atomicOr(err, DEV_ERR_WATCHDOG);
          ok = 0u;
          break;
        }
        __nanosleep(64);
      }
    }
    __threadfence();
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0u;
}

template <int MINB>
__global__ void __launch_bounds__(RP_THREADS, MINB) rl_persistent_kernel(const RpParams p) {
  __shared__ uint32_t s_wtot[RP_THREADS / 32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x / p.ctas_per_group, c = blockIdx.x - g * p.ctas_per_group;
  const uint32_t C = (uint32_t)p.ctas_per_group;
  const int64_t gthreads = (int64_t)C * RP_THREADS;
  const int64_t gtid = (int64_t)c * RP_THREADS + tid;
  unsigned long long* tab = p.tables + (size_t)g * ((size_t)p.cap_mask + 1);
  uint32_t* slot_of = p.slot_of + (size_t)g * p.n_pad;
  uint8_t* fbytes = p.fbytes + (size_t)g * p.n_pad;
  uint32_t* cta_count = p.cta_count + (size_t)g * C;
  uint32_t* bar = p.bars + (size_t)g * 32;
  const int64_t S = p.num_seeds;
  const uint32_t mask = p.cap_mask;

  for (int b = g; b < p.num_trees; b += p.groups) {
    int64_t n = p.lens[b];
    n = n < 0 ? 0 : (n > p.n_max ? p.n_max : n);
    const int64_t* src = p.samples + (int64_t)b * p.stride;
    int64_t* local = p.local + (int64_t)b * p.stride;
    int64_t* nodes = p.nodes + (int64_t)b * p.stride;

    // ---- clear the group's table (16-byte stores; the table stays in the L2) -----------------------------------
    {
      ulonglong2* t2 = reinterpret_cast<ulonglong2*>(tab);
      const int64_t n2 = ((int64_t)mask + 1) >> 1;
      for (int64_t q = gtid; q < n2; q += gthreads) t2[q] = make_ulonglong2(RP_EMPTY, RP_EMPTY);
    }
    if (!group_sync(bar, C, p.err)) return;

    // ---- insert: four ids per thread in flight ------------------------------------------------------------------
    for (int64_t i0 = gtid; i0 < n; i0 += 4 * gthreads) {
      int64_t key[4];
      unsigned long long want[4], old[4];
      uint32_t h[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * gthreads;
        key[u] = i < n ? __ldg(src + i) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * gthreads;
        old[u] = RP_EMPTY;
        want[u] = 0ull;
        h[u] = RL_NOSLOT;
        if (i >= n) continue;
        if ((uint64_t)key[u] >= 0xFFFFFFFFull) {
          atomicOr(p.err, DEV_ERR_INDEX);
          continue;
        }
        const uint32_t prio = i < S ? (uint32_t)(S - 1 - i) : (uint32_t)i;
        want[u] = ((unsigned long long)(uint32_t)key[u] << 32) | prio;
        h[u] = (((uint32_t)key[u] * 0x9E3779B1u) >> p.hash_shift) & mask;
        old[u] = atomicCAS(tab + h[u], RP_EMPTY, want[u]);  // the common case (a new id, a free slot): one atomic
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * gthreads;
        if (i >= n) continue;
        if (h[u] != RL_NOSLOT) {
          unsigned long long o = old[u];
          // double hashing (odd step: every slot of the power-of-two table is visited): a group waits at its barrier
          // for the LONGEST probe sequence of the tree, and linear probing's clusters made that 30 % of the kernel
          const uint32_t step = (((uint32_t)key[u] * 0x85EBCA6Bu) >> 7) | 1u;
          while (o != RP_EMPTY) {                                  // slot taken
            if ((uint32_t)(o >> 32) == (uint32_t)key[u]) {          // by the same id: the smaller priority wins
              atomicMin(tab + h[u], want[u]);
              break;
            }
            h[u] = (h[u] + step) & mask;                            // by another id: next slot of the id's sequence
            o = atomicCAS(tab + h[u], RP_EMPTY, want[u]);
          }
        }
        slot_of[i] = h[u];
      }
    }
    if (!group_sync(bar, C, p.err)) return;

    // ---- count: per-position state, local[] of everything that needs no rank, flags per CTA chunk ---------------
    // CTA c owns the contiguous positions [c0, c1); a thread owns RP_ITEMS consecutive ones (one uint4 of slot_of)
    int64_t chunk = (n + C - 1) / C;
    chunk = (chunk + RP_TILE - 1) / RP_TILE * RP_TILE;
    const int64_t c0 = min(n, (int64_t)c * chunk), c1 = min(n, c0 + chunk);
    uint32_t my_flags = 0;
    for (int64_t t0 = c0; t0 < c1; t0 += RP_TILE) {
      const int64_t ibase = t0 + (int64_t)tid * RP_ITEMS;
      if (ibase >= c1) continue;
      const uint4 sv = __ldcg(reinterpret_cast<const uint4*>(slot_of + ibase));
      const uint32_t sl[4] = {sv.x, sv.y, sv.z, sv.w};
      unsigned long long e[4];
#pragma unroll
      for (int u = 0; u < RP_ITEMS; ++u) e[u] = (ibase + u < c1 && sl[u] != RL_NOSLOT) ? __ldcg(tab + sl[u]) : 0ull;
      uint32_t st4 = 0;
#pragma unroll
      for (int u = 0; u < RP_ITEMS; ++u) {
        const int64_t i = ibase + u;
        if (i >= c1) continue;
        uint32_t st = 0;
        if (sl[u] == RL_NOSLOT) {                 // id outside the 32-bit range (already reported)
          st = i < S ? RP_F_NODE : 0u;
          local[i] = -1;
        } else {
          const uint32_t pr = (uint32_t)e[u];
          if (pr < (uint32_t)S) {                 // a seed carries this id: the map points at its LAST seed slot (:26)
            st_cs_i64(local + i, S - 1 - (int64_t)pr);
            if (i < S) st = RP_F_NODE;            // every seed is kept (:25); its rank is its own index
          } else if (pr == (uint32_t)i) {
            st = RP_F_NODE | RP_F_PUBLISH;        // first occurrence of an id no seed carries (:36-39)
          } else {
            st = RP_F_PENDING;
          }
        }
        my_flags += st & RP_F_NODE;
        st4 |= st << (8 * u);
      }
      *reinterpret_cast<uint32_t*>(fbytes + ibase) = st4;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_flags += __shfl_xor_sync(0xffffffffu, my_flags, o);
    if (lane == 0) s_wtot[warp] = my_flags;
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < RP_THREADS / 32; ++w) tot += s_wtot[w];
      cta_count[c] = tot;
    }
    if (!group_sync(bar, C, p.err)) return;

    // ---- assign: ranks in position order, node list, ranks published in the table -----------------------------
    {
      uint32_t before = 0, total = 0;
      for (uint32_t q = lane; q < C; q += 32) {
        const uint32_t v = __ldcg(cta_count + q);
        total += v;
        if (q < (uint32_t)c) before += v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
      }
      if (c == 0 && tid == 0) p.nodes_len[b] = (int64_t)total;
      if (tid == 0) s_carry = before;
    }
    __syncthreads();
    for (int64_t t0 = c0; t0 < c1; t0 += RP_TILE) {
      const int64_t ibase = t0 + (int64_t)tid * RP_ITEMS;
      const uint32_t st4 = ibase < c1 ? __ldcg(reinterpret_cast<const uint32_t*>(fbytes + ibase)) : 0u;
      const uint32_t cnt = __popc(st4 & 0x01010101u);
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) s_wtot[warp] = incl;
      __syncthreads();
      uint32_t excl = incl - cnt, tile_total = 0;
#pragma unroll
      for (int w = 0; w < RP_THREADS / 32; ++w) {
        const uint32_t v = s_wtot[w];
        if (w < warp) excl += v;
        tile_total += v;
      }
      uint32_t r = s_carry + excl;
      if (cnt) {
        const uint4 sv = __ldcg(reinterpret_cast<const uint4*>(slot_of + ibase));
        const uint32_t sl[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int u = 0; u < RP_ITEMS; ++u) {
          const uint32_t st = (st4 >> (8 * u)) & 0xffu;
          if (!(st & RP_F_NODE)) continue;
          const int64_t i = ibase + u;
          const int64_t key = __ldg(src + i);
          st_cs_i64(nodes + r, key);
          if (st & RP_F_PUBLISH) {
            st_cs_i64(local + i, (int64_t)r);
            __stcg(tab + sl[u], ((unsigned long long)(uint32_t)key << 32) | RP_RANKED | r);
          }
          ++r;
        }
      }
      __syncthreads();
      if (tid == 0) s_carry += tile_total;
      __syncthreads();
    }
    if (!group_sync(bar, C, p.err)) return;

    // ---- lookup: later occurrences of non-seed ids take the rank their first occurrence published -----------
    for (int64_t i4 = gtid * 4; i4 < n; i4 += 4 * gthreads) {
      const uint32_t st4 = __ldcg(reinterpret_cast<const uint32_t*>(fbytes + i4));
      if (!(st4 & 0x04040404u)) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (((st4 >> (8 * u)) & RP_F_PENDING) && i4 + u < n)
          st_cs_i64(local + i4 + u, (int64_t)((uint32_t)__ldcg(tab + __ldcg(slot_of + i4 + u)) & 0x7fffffffu));
    }
    if (!group_sync(bar, C, p.err)) return;   // the next tree's clear must not overtake these reads
  }
}

struct RpLayout {
  uint32_t slots;
  int log2_slots, groups;
  int64_t n_pad;
  size_t off_bars, off_cta, off_tables, off_slot_of, off_fbytes, total;
};

bool rp_layout(int64_t num_trees, int64_t n_max, RpLayout& L) {
  if (num_trees <= 0 || n_max < 0 || n_max >= ((int64_t)1 << 30)) return false;
  uint64_t slots = 1024;
  while (slots < 2 * (uint64_t)n_max + 2) slots <<= 1;   // load <= 0.5 in the worst case, about 0.3 for sampled trees
  L.slots = (uint32_t)slots;
  L.log2_slots = 0;
  while ((1ull << L.log2_slots) < slots) ++L.log2_slots;
  L.n_pad = (n_max + RP_ITEMS - 1) / RP_ITEMS * RP_ITEMS + RP_ITEMS;
  // groups: as many trees in flight as keep tables + slot maps + state bytes + ids inside the L2 budget
  const size_t per_group = slots * 8 + (size_t)L.n_pad * 13;
  const char* e = getenv("TCHGEO_RELABEL_GROUPS");
  int64_t groups = e ? atoi(e) : (int64_t)(((size_t)72 << 20) / std::max<size_t>(per_group, 1));
  groups = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(groups, num_trees), RP_MAX_GROUPS));
  L.groups = (int)groups;
  size_t o = 0;
  L.off_bars = o; o += rl_align((size_t)RP_MAX_GROUPS * 128);
  L.off_cta = o; o += rl_align((size_t)RP_MAX_CTAS * 4);
  L.off_tables = o; o += rl_align((size_t)groups * slots * 8);
  L.off_slot_of = o; o += rl_align((size_t)groups * L.n_pad * 4);
  L.off_fbytes = o; o += rl_align((size_t)groups * L.n_pad);
  L.total = o + 256;
  return true;
}

// -> cudaErrorNotSupported when the device cannot launch cooperatively (the caller then runs the wave form)
cudaError_t rp_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees, int64_t num_seeds,
                       int64_t n_max, int64_t* nodes, int64_t* local, int64_t* nodes_len, char* ws, const RpLayout& L,
                       uint32_t* err, cudaStream_t stream) {
  // tuning knob: CTAs per SM the register budget is set for (2: 64 registers, 3: 40)
  static const int minb = [] { const char* e = getenv("TCHGEO_RELABEL_MINB"); return (e && atoi(e) == 3) ? 3 : 2; }();
  const void* kernel = minb == 3 ? (const void*)rl_persistent_kernel<3> : (const void*)rl_persistent_kernel<2>;
  static int resident[64] = {};  // co-resident CTAs of the kernel per device (0 = not queried, -1 = no cooperative launch)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorNotSupported;
  if (resident[dev] == 0) {
    int coop = 0, sms = 0, per_sm = 0;
    e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) {
      e = minb == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rl_persistent_kernel<3>, RP_THREADS, 0)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rl_persistent_kernel<2>, RP_THREADS, 0);
    }
    if (e != cudaSuccess) return e;
    resident[dev] = (coop && sms * per_sm > 0) ? sms * per_sm : -1;
  }
  if (resident[dev] < 0) return cudaErrorNotSupported;
  RpParams p;
  p.samples = samples; p.stride = stride; p.lens = lens; p.nodes = nodes; p.local = local; p.nodes_len = nodes_len;
  p.num_seeds = num_seeds; p.n_max = n_max; p.n_pad = L.n_pad; p.num_trees = (int32_t)num_trees;
  p.groups = std::min(L.groups, resident[dev]);
  p.ctas_per_group = std::min(resident[dev], RP_MAX_CTAS) / p.groups;
  p.cap_mask = L.slots - 1; p.hash_shift = 32 - L.log2_slots;
  p.tables = (unsigned long long*)(ws + L.off_tables);
  p.slot_of = (uint32_t*)(ws + L.off_slot_of);
  p.fbytes = (uint8_t*)(ws + L.off_fbytes);
  p.cta_count = (uint32_t*)(ws + L.off_cta);
  p.bars = (uint32_t*)(ws + L.off_bars);
  p.err = err;
  e = cudaMemsetAsync(ws + L.off_bars, 0, (size_t)RP_MAX_GROUPS * 128, stream);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel(kernel, dim3((unsigned)(p.groups * p.ctas_per_group)),
                                     dim3(RP_THREADS), args, 0, stream);
}

__global__ void rl_set_len_kernel(int64_t* p, int64_t v) { *p = v; }

struct RlLayout {
  uint32_t slots;        // per tree, power of two > n_max
  int log2_slots;
  int wave;              // trees per wave
  int tiles_per_tree;
  size_t table_bytes;    // per tree: keys + priorities
  // byte offsets into the workspace
  size_t off_ctrl;       // [0] err (u32), [8] single-tree length (i64), [16] single-tree nodes_len (i64), [64..] one ticket per wave
  size_t off_tables;     // wave tables, then the look-back status of the wave: one 0xFF memset covers both
  size_t fill_bytes;
  size_t off_rank, off_slot_of, total;
  int num_waves;
};

bool rl_layout(int64_t num_trees, int64_t n_max, bool k32, RlLayout& L) {
  if (num_trees <= 0 || n_max < 0 || n_max >= ((int64_t)1 << 31)) return false;
  uint64_t slots = 1024;
  while (slots < (uint64_t)n_max + (uint64_t)n_max / 8 + 2) slots <<= 1;  // load factor <= 0.89, typically about half that
  L.slots = (uint32_t)slots;
  L.log2_slots = 0;
  while ((1ull << L.log2_slots) < slots) ++L.log2_slots;
  L.table_bytes = slots * (k32 ? Table<true>::slot_bytes : Table<false>::slot_bytes);
  L.tiles_per_tree = (int)std::max<int64_t>(1, (n_max + RL_TILE - 1) / RL_TILE);
  // per-tree L2 footprint: table + ranks + slot_of + the ids themselves
  const size_t per_tree = L.table_bytes + slots * 4 + (size_t)n_max * 12;
  const char* e = getenv("TCHGEO_RELABEL_WAVE_MB");
  const size_t budget = (size_t)std::max(1, e ? atoi(e) : 64) << 20;
  int64_t wave = (int64_t)(budget / std::max<size_t>(per_tree, 1));
  wave = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(wave, num_trees), 16384));
  L.wave = (int)wave;
  L.num_waves = (int)((num_trees + wave - 1) / wave);
  size_t o = 0;
  L.off_ctrl = o; o += rl_align(64 + (size_t)L.num_waves * 4);
  L.off_tables = o;
  L.fill_bytes = rl_align((size_t)wave * L.table_bytes) + rl_align((size_t)wave * L.tiles_per_tree * 8);
  o += L.fill_bytes;
  L.off_rank = o; o += rl_align((size_t)wave * slots * 4);
  L.off_slot_of = o; o += rl_align((size_t)wave * (size_t)std::max<int64_t>(n_max, 1) * 4);
  L.total = o + 256;
  return true;
}

// Enqueues the whole stage; device-side errors are OR-ed into *err (DEVICE, not cleared here).
tchgeo_status rl_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees, int64_t num_seeds,
                         int64_t n_max, bool k32, int64_t* nodes, int64_t* local, int64_t* nodes_len, char* ws,
                         const RlLayout& L, uint32_t* err, cudaStream_t stream) {
  RlParams p;
  p.samples = samples; p.stride = stride; p.lens = lens; p.nodes = nodes; p.local = local; p.nodes_len = nodes_len;
  p.num_seeds = num_seeds; p.n_max = n_max; p.tiles_per_tree = L.tiles_per_tree;
  p.cap_mask = L.slots - 1; p.hash_shift = 32 - L.log2_slots;
  p.tables = ws + L.off_tables; p.table_bytes = L.table_bytes;
  p.rank = (uint32_t*)(ws + L.off_rank); p.slot_of = (uint32_t*)(ws + L.off_slot_of);
  p.status = (uint64_t*)(ws + L.off_tables + rl_align((size_t)L.wave * L.table_bytes));
  p.err = err;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_ctrl + 64, 0, (size_t)L.num_waves * 4, stream));
  for (int w = 0; w < L.num_waves; ++w) {
    const int b0 = w * L.wave;
    const int nb = (int)std::min<int64_t>(L.wave, num_trees - b0);
    p.b0 = b0;
    p.ticket = (uint32_t*)(ws + L.off_ctrl + 64) + w;
    // empty keys, maximal priorities and "not published" status words are all 0xFF bytes
    TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_tables, 0xFF, L.fill_bytes, stream));
    const dim3 grid((unsigned)L.tiles_per_tree, (unsigned)nb);
    if (k32) rl_insert_kernel<true><<<grid, RL_THREADS, 0, stream>>>(p);
    else rl_insert_kernel<false><<<grid, RL_THREADS, 0, stream>>>(p);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    const unsigned tiles = (unsigned)L.tiles_per_tree * (unsigned)nb;
    if (k32) rl_compact_kernel<true><<<tiles, RL_THREADS, 0, stream>>>(p);
    else rl_compact_kernel<false><<<tiles, RL_THREADS, 0, stream>>>(p);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
    rl_lookup_kernel<<<grid, RL_THREADS, 0, stream>>>(p);
    TCHGEO_CUDA_CHECK(cudaGetLastError());
  }
  return TCHGEO_OK;
}

}  // namespace

// used by the sampling plan (neighbor_sampling.cu): size and enqueue the stage for one node type.  k32 (ids < 2^32-1): the
// persistent form, with the wave form as its fallback on devices without cooperative launch (TCHGEO_RELABEL_WAVES=1
// forces it); any i64 id: the wave form with 64-bit keys.
static bool use_persistent(bool k32) {
  const char* e = getenv("TCHGEO_RELABEL_WAVES");
  return k32 && !(e && atoi(e) != 0);
}
size_t relabel_workspace_bytes(int64_t num_trees, int64_t n_max, bool k32) {
  RlLayout L;
  if (!rl_layout(num_trees, n_max, k32, L)) return 0;
  RpLayout P;
  if (use_persistent(k32) && rp_layout(num_trees, n_max, P)) return std::max(L.total, P.total);
  return L.total;
}
int relabel_launches(int64_t num_trees, int64_t n_max, bool k32) {
  if (use_persistent(k32)) return 1;
  RlLayout L;
  return rl_layout(num_trees, n_max, k32, L) ? 3 * L.num_waves : 0;
}
tchgeo_status relabel_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees,
                              int64_t num_seeds, int64_t n_max, bool k32, int64_t* nodes, int64_t* local,
                              int64_t* nodes_len, void* workspace, size_t workspace_bytes, uint32_t* err,
                              cudaStream_t stream) {
  RlLayout L;
  TCHGEO_REQUIRE(rl_layout(num_trees, n_max, k32, L), "relabel: tree too large");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= relabel_workspace_bytes(num_trees, n_max, k32),
                 "relabel: workspace too small (need %zu bytes)", relabel_workspace_bytes(num_trees, n_max, k32));
  RpLayout P;
  if (use_persistent(k32) && rp_layout(num_trees, n_max, P)) {
    const cudaError_t e = rp_enqueue(samples, stride, lens, num_trees, num_seeds, n_max, nodes, local, nodes_len,
                                     (char*)workspace, P, err, stream);
    if (e == cudaSuccess) return TCHGEO_OK;
    if (e != cudaErrorNotSupported) TCHGEO_CUDA_CHECK(e);
    (void)cudaGetLastError();
  }
  return rl_enqueue(samples, stride, lens, num_trees, num_seeds, n_max, k32, nodes, local, nodes_len, (char*)workspace, L,
                    err, stream);
}

}  // namespace tchgeo

using namespace tchgeo;

extern "C" size_t tchgeo_unique_relabel_batched_workspace_bytes(int64_t num_batches, int64_t n_max, int32_t key32) {
  return relabel_workspace_bytes(num_batches, n_max, key32 != 0);
}

extern "C" tchgeo_status tchgeo_unique_relabel_batched(const int64_t* samples, int64_t stride, const int64_t* lens,
                                                       int64_t num_batches, int64_t num_seeds, int64_t n_max,
                                                       int32_t key32, int64_t* nodes, int64_t* local, int64_t* nodes_len,
                                                       void* workspace, size_t workspace_bytes, int32_t* err_word,
                                                       tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_batches >= 0 && stride >= 0 && n_max >= 0 && n_max <= stride, "bad relabel geometry");
  TCHGEO_REQUIRE(num_seeds >= 0 && num_seeds <= n_max, "num_seeds out of range");
  if (num_batches == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(samples && lens && nodes && local && nodes_len && err_word, "NULL pointer");
  return relabel_enqueue(samples, stride, lens, num_batches, num_seeds, n_max, key32 != 0, nodes, local, nodes_len, workspace,
                         workspace_bytes, (uint32_t*)err_word, (cudaStream_t)stream_);
}

extern "C" size_t tchgeo_unique_relabel_workspace_bytes(int64_t n) {
  if (n < 0 || n >= ((int64_t)1 << 30)) return 0;
  return relabel_workspace_bytes(1, n, false);
}

// One tree (the B = 1 case of the batched stage, any i64 ids), synchronous: returns the number of nodes.
extern "C" tchgeo_status tchgeo_unique_relabel(const int64_t* samples, int64_t n, int64_t num_seeds, int64_t* nodes,
                                               int64_t* local, int64_t* num_nodes, void* workspace,
                                               size_t workspace_bytes, tchgeo_stream stream_) {
  TCHGEO_REQUIRE(n >= 0 && n < ((int64_t)1 << 30), "n out of range");
  TCHGEO_REQUIRE(num_seeds >= 0 && num_seeds <= n, "num_seeds out of range");
  if (n == 0) {
    if (num_nodes) *num_nodes = 0;
    return TCHGEO_OK;
  }
  TCHGEO_REQUIRE(samples && nodes && local && workspace, "NULL pointer");
  RlLayout L;
  TCHGEO_REQUIRE(rl_layout(1, n, false, L), "n out of range");
  if (workspace_bytes < L.total) {
    set_last_error("workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return TCHGEO_ERR_CAPACITY;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  uint32_t* d_err = (uint32_t*)(ws + L.off_ctrl);
  int64_t* d_len = (int64_t*)(ws + L.off_ctrl + 8);
  int64_t* d_nodes_len = (int64_t*)(ws + L.off_ctrl + 16);
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_ctrl, 0, 64, stream));
  rl_set_len_kernel<<<1, 1, 0, stream>>>(d_len, n);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  tchgeo_status st = rl_enqueue(samples, n, d_len, 1, num_seeds, n, false, nodes, local, d_nodes_len, ws, L, d_err, stream);
  if (st != TCHGEO_OK) return st;
  int64_t h[3] = {0, 0, 0};
  TCHGEO_CUDA_CHECK(cudaMemcpyAsync(h, ws + L.off_ctrl, 24, cudaMemcpyDeviceToHost, stream));
  TCHGEO_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (num_nodes) *num_nodes = h[2];
  return status_from_dev_err((uint32_t)h[0]);
}
