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
#include <cstdio>
#include <cstdlib>
#include <vector>

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
// The grid is G groups of C co-resident CTAs; a group owns TWO hash tables and walks its trees two at a time:
//     insert(a) insert(b) compact(a) compact(b) clear(a) clear(b) insert(a') ...
// Only two things make a phase wait -- "every CTA of the group has finished the previous phase ON THE SAME TABLE" --
// and between that arrival and the wait lies a whole phase of the OTHER tree, so the barrier latency and the spread of
// the CTAs' finishing times are hidden (split arrive / wait on monotonic counters; the first version had five blocking
// barriers per tree and spent 52 % of its warp time in them).
//   insert   one table slot is ONE 64-bit word (key << 32 | priority): a new id is a single atomicCAS, a second
//            occurrence adds one 64-bit atomicMin (equal keys: the smaller priority wins); double hashing (a group waits
//            for the LONGEST probe sequence of the tree); eight ids per thread in flight
//   compact  flags, block scan, decoupled look-back over the tree's tiles (tiles are taken in order, so every tile a
//            look-back waits for is running or done), node list, local ids.  The first occurrence of an id overwrites its
//            priority with (1 << 31 | rank); a later occurrence spins on that bit -- its winner sits at a smaller
//            position, i.e. in an earlier tile or in this one -- so no separate lookup pass and no state bytes
// At any moment only 2G tables, slot maps and trees are live, so table accesses and re-reads of the ids are L2 hits and
// DRAM sees 8 B read + (8 B local + <= 8 B nodes) written per id.  Everything CTAs hand to each other is read with ld.cg /
// ld.relaxed.gpu: the L1 is not coherent and the same addresses are reused tree after tree.
// The co-residency the waits rely on is what cudaLaunchCooperativeKernel guarantees (two plans on two streams would
// otherwise be able to starve each other's CTAs).
// =================================================================================================
constexpr int RP_THREADS = 512;
constexpr int RP_ITEMS = 4;                       // consecutive positions per thread in the compact phase
constexpr int RP_TILE = RP_THREADS * RP_ITEMS;
constexpr int RP_INS = 8;                         // ids per thread in flight in the insert phase
constexpr int RP_MAX_GROUPS = 16;
constexpr int RP_MAX_CTAS = 4096;
constexpr unsigned long long RP_EMPTY = ~0ull;
constexpr uint32_t RP_RANKED = 0x80000000u;
constexpr uint64_t RP_EPOCH_SHIFT = 40;           // look-back word: flag (2) | epoch (22) | value (40)
constexpr uint64_t RP_VAL_MASK = (1ull << RP_EPOCH_SHIFT) - 1;
enum { RP_EV_CLEAR = 0, RP_EV_INSERT = 1, RP_EV_COMPACT = 2 };

struct RpParams {
  const int64_t* samples;
  int64_t stride;
  const int64_t* lens;
  int64_t* nodes;
  int64_t* local;
  int64_t* nodes_len;
  int64_t num_seeds, n_max, n_pad;   // n_pad: row pitch of the slot maps (multiple of RP_ITEMS)
  int32_t num_trees, groups, ctas_per_group, tiles_per_tree;
  uint32_t cap_mask;
  int32_t hash_shift;
  unsigned long long* tables;        // [groups, 2, slots]
  uint32_t* slot_of;                 // [groups, 2, n_pad]
  unsigned long long* status;        // [groups, 2, tiles_per_tree] look-back words, tagged with the tree's epoch
  uint32_t* events;                  // [groups, 32]: monotonic arrival counters [kind * 2 + table] (zero-initialised)
  uint32_t* err;
  unsigned long long* trace;         // debug (TCHGEO_RELABEL_TRACE): [CTAs, RP_TRACE] globaltimer stamps, or NULL
};
constexpr int RP_TRACE = 128;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// this CTA has finished a phase: everything it wrote becomes visible before the counter moves
__device__ __forceinline__ void rp_arrive(uint32_t* ev) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ev, 1u);
  }
}
// every CTA of the group has arrived `target` times.  Returns false when the watchdog trips.
__device__ __forceinline__ bool rp_wait(const uint32_t* ev, uint32_t target, uint32_t* err) {
  __shared__ uint32_t s_ok;
  if (threadIdx.x == 0) {
    uint32_t ok = 1u, spins = 0;
    while ((int32_t)(ld_acquire_u32(ev) - target) < 0) {
      if (++spins > (1u << 26)) {
        atomicOr(err, DEV_ERR_WATCHDOG);
        ok = 0u;
        break;
      }
      __nanosleep(40);
    }
    s_ok = ok;
  }
  __syncthreads();
  const bool ok = s_ok != 0u;
  __syncthreads();   // s_ok may be rewritten by the next wait
  return ok;
}

struct RpTree {
  const int64_t* src;
  int64_t* local;
  int64_t* nodes;
  int64_t n;
  int b;
};

template <int MINB>
__global__ void __launch_bounds__(RP_THREADS, MINB) rl_persistent_kernel(const RpParams p) {
  __shared__ uint32_t s_wtot[RP_THREADS / 32];
  __shared__ int64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x / p.ctas_per_group, c = blockIdx.x - g * p.ctas_per_group;
  const uint32_t C = (uint32_t)p.ctas_per_group;
  const size_t slots = (size_t)p.cap_mask + 1;
  unsigned long long* tab2 = p.tables + (size_t)g * 2 * slots;
  uint32_t* slot2 = p.slot_of + (size_t)g * 2 * p.n_pad;
  unsigned long long* status2 = p.status + (size_t)g * 2 * p.tiles_per_tree;
  uint32_t* ev = p.events + (size_t)g * 32;
  const int64_t S = p.num_seeds;
  const uint32_t mask = p.cap_mask;
  const int M = p.num_trees > g ? (p.num_trees - g + p.groups - 1) / p.groups : 0;  // trees of this group: g, g+G, ...

  int n_stamp = 0;
  auto stamp = [&]() {   // debug: when did this CTA pass this point (after a wait / after a phase body)
    if (p.trace && tid == 0 && n_stamp < RP_TRACE) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
      p.trace[(size_t)blockIdx.x * RP_TRACE + n_stamp++] = ns;
    }
  };
  auto tree_of = [&](int m) {
    RpTree t;
    t.b = g + m * p.groups;
    int64_t n = p.lens[t.b];
    t.n = n < 0 ? 0 : (n > p.n_max ? p.n_max : n);
    t.src = p.samples + (int64_t)t.b * p.stride;
    t.local = p.local + (int64_t)t.b * p.stride;
    t.nodes = p.nodes + (int64_t)t.b * p.stride;
    return t;
  };

  auto clear = [&](int tb) {   // 16-byte stores; the table stays in the L2
    ulonglong2* t2 = reinterpret_cast<ulonglong2*>(tab2 + (size_t)tb * slots);
    const int64_t n2 = (int64_t)(slots >> 1);
    for (int64_t q = (int64_t)c * RP_THREADS + tid; q < n2; q += (int64_t)C * RP_THREADS)
      t2[q] = make_ulonglong2(RP_EMPTY, RP_EMPTY);
  };

  auto insert = [&](int m) {
    const RpTree t = tree_of(m);
    unsigned long long* tab = tab2 + (size_t)(m & 1) * slots;
    uint32_t* slot_of = slot2 + (size_t)(m & 1) * p.n_pad;
    constexpr int64_t BLK = (int64_t)RP_THREADS * RP_INS;
    for (int64_t i0 = (int64_t)c * BLK; i0 < t.n; i0 += (int64_t)C * BLK) {
      // (32-bit keys and no `want` array: eight inserts in flight must fit the 64-register budget)
      uint32_t key[RP_INS], h[RP_INS];
      unsigned long long old[RP_INS];
#pragma unroll
      for (int u = 0; u < RP_INS; ++u) {
        const int64_t i = i0 + u * RP_THREADS + tid;
        h[u] = RL_NOSLOT;
        key[u] = 0xFFFFFFFFu;
        if (i < t.n) {
          const int64_t k64 = __ldg(t.src + i);
          if ((uint64_t)k64 >= 0xFFFFFFFFull) atomicOr(p.err, DEV_ERR_INDEX);
          else key[u] = (uint32_t)k64;
        }
      }
      auto want_of = [&](int u) {
        const int64_t i = i0 + u * RP_THREADS + tid;
        const uint32_t prio = i < S ? (uint32_t)(S - 1 - i) : (uint32_t)i;
        return ((unsigned long long)key[u] << 32) | prio;
      };
#pragma unroll
      for (int u = 0; u < RP_INS; ++u) {
        old[u] = RP_EMPTY;
        if (key[u] == 0xFFFFFFFFu) continue;
        h[u] = ((key[u] * 0x9E3779B1u) >> p.hash_shift) & mask;
        old[u] = atomicCAS(tab + h[u], RP_EMPTY, want_of(u));  // the common case (a new id, a free slot): one atomic
      }
      // collisions are resolved in ROUNDS: every unresolved id of the thread issues its next probe before any result is
      // looked at, so a thread pays max (not sum) of its ids' probe lengths in L2 round trips
      bool again = true;
      while (again) {
        again = false;
#pragma unroll
        for (int u = 0; u < RP_INS; ++u) {
          if (h[u] == RL_NOSLOT || old[u] == RP_EMPTY) continue;          // resolved (or no id)
          if ((uint32_t)(old[u] >> 32) == key[u]) {                          // slot holds the same id: the smaller priority wins
            atomicMin(tab + h[u], want_of(u));
            old[u] = RP_EMPTY;
          } else {                                                           // another id: next slot of the id's sequence
            const uint32_t step = ((key[u] * 0x85EBCA6Bu) >> 7) | 1u;        // odd: visits every slot (double hashing)
            h[u] = (h[u] + step) & mask;
            old[u] = atomicCAS(tab + h[u], RP_EMPTY, want_of(u));
            again = true;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RP_INS; ++u) {
        const int64_t i = i0 + u * RP_THREADS + tid;
        if (i < t.n) slot_of[i] = h[u];
      }
    }
  };

  auto compact = [&](int m) {
    const RpTree t = tree_of(m);
    unsigned long long* tab = tab2 + (size_t)(m & 1) * slots;
    const uint32_t* slot_of = slot2 + (size_t)(m & 1) * p.n_pad;
    unsigned long long* status = status2 + (size_t)(m & 1) * p.tiles_per_tree;
    const uint64_t epoch = (uint64_t)(t.b + 1) << RP_EPOCH_SHIFT;   // words of earlier trees read as "not published"
    const int64_t ntiles = t.n > 0 ? (t.n + RP_TILE - 1) / RP_TILE : 1;
    for (int64_t tile = c; tile < ntiles; tile += C) {   // in increasing order: look-backs only wait for smaller tiles
      const int64_t ibase = tile * RP_TILE + (int64_t)tid * RP_ITEMS;
      uint32_t sl[RP_ITEMS] = {RL_NOSLOT, RL_NOSLOT, RL_NOSLOT, RL_NOSLOT};
      unsigned long long e[RP_ITEMS];
      if (ibase < t.n) {
        const uint4 sv = __ldcg(reinterpret_cast<const uint4*>(slot_of + ibase));
        sl[0] = sv.x; sl[1] = sv.y; sl[2] = sv.z; sl[3] = sv.w;
      }
#pragma unroll
      for (int u = 0; u < RP_ITEMS; ++u) e[u] = (ibase + u < t.n && sl[u] != RL_NOSLOT) ? __ldcg(tab + sl[u]) : 0ull;
      uint32_t node = 0, pend = 0;   // bit u: position ibase+u emits a node / waits for its winner's rank
      uint32_t cnt = 0;
#pragma unroll
      for (int u = 0; u < RP_ITEMS; ++u) {
        const int64_t i = ibase + u;
        if (i >= t.n) continue;
        if (sl[u] == RL_NOSLOT) {                 // id outside the 32-bit range (already reported)
          if (i < S) node |= 1u << u;
          t.local[i] = -1;
          continue;
        }
        const uint32_t pr = (uint32_t)e[u];
        if (pr < (uint32_t)S) {                   // a seed carries this id: the map points at its LAST seed slot (:26)
          st_cs_i64(t.local + i, S - 1 - (int64_t)pr);
          if (i < S) node |= 1u << u;             // every seed is kept (:25); its rank is its own index
        } else if (pr == (uint32_t)i) {
          node |= 1u << u;                        // first occurrence of an id no seed carries (:36-39)
        } else {
          pend |= 1u << u;
        }
      }
      cnt = __popc(node);
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
      for (int w = 0; w < RP_THREADS / 32; ++w) {
        const uint32_t v = s_wtot[w];
        if (w < warp) excl += v;
        total += v;
      }
      if (tid == 0) st_relaxed_u64((uint64_t*)status + tile, ((tile == 0 ? 2ull : 1ull) << 62) | epoch | (uint64_t)total);
      if (warp == 0) {
        int64_t before = 0;
        if (tile > 0) {
          int64_t j = tile - 1;
          uint32_t spins = 0;
          while (true) {
            const int64_t idx = j - lane;
            uint64_t v = idx >= 0 ? ld_relaxed_u64((const uint64_t*)status + idx) : (2ull << 62) | epoch;
            if ((v & (((1ull << 22) - 1) << RP_EPOCH_SHIFT)) != epoch) v = 0;   // an earlier tree's word
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
              __nanosleep(20);
              continue;
            }
            int64_t val = lane <= first_incl ? (int64_t)(v & RP_VAL_MASK) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            before += val;
            if (first_incl < 32) break;
            j -= 32;
          }
          if (lane == 0) st_relaxed_u64((uint64_t*)status + tile, (2ull << 62) | epoch | (uint64_t)(before + total));
        }
        if (lane == 0) s_excl = before;
      }
      __syncthreads();
      const int64_t tile_excl = s_excl;
      if (tid == 0 && tile == ntiles - 1) p.nodes_len[t.b] = tile_excl + total;
      uint32_t r = (uint32_t)tile_excl + excl;
#pragma unroll
      for (int u = 0; u < RP_ITEMS; ++u) {
        if (!((node >> u) & 1u)) continue;
        const int64_t i = ibase + u;
        const int64_t key = __ldg(t.src + i);
        st_cs_i64(t.nodes + r, key);
        if (i >= S && sl[u] != RL_NOSLOT) {     // the occurrence the map points at publishes its rank in place
          st_cs_i64(t.local + i, (int64_t)r);
          st_relaxed_u64((uint64_t*)tab + sl[u], ((uint64_t)(uint32_t)key << 32) | RP_RANKED | r);
        }
        ++r;
      }
      __syncthreads();   // this tile's ranks are on their way before anyone of the CTA starts to wait for ranks
      if (pend) {
#pragma unroll
        for (int u = 0; u < RP_ITEMS; ++u) {
          if (!((pend >> u) & 1u)) continue;
          uint32_t lo = (uint32_t)e[u], spins = 0;
          while (!(lo & RP_RANKED)) {             // the winner sits at a smaller position: an earlier tile, or this one
            lo = (uint32_t)ld_relaxed_u64((const uint64_t*)tab + sl[u]);
            if (++spins > (1u << 24)) {
              atomicOr(p.err, DEV_ERR_WATCHDOG);
              break;
            }
          }
          st_cs_i64(t.local + ibase + u, (int64_t)(lo & 0x7fffffffu));
        }
      }
    }
  };

  // ---- the group's schedule: two trees in flight, every wait separated from its arrival by a phase of the other tree ----
  stamp();
  clear(0);
  rp_arrive(ev + RP_EV_CLEAR * 2 + 0);
  clear(1);
  rp_arrive(ev + RP_EV_CLEAR * 2 + 1);
  stamp();
  for (int pr = 0; 2 * pr < M; ++pr) {
    const uint32_t target = C * (uint32_t)(pr + 1);
    for (int tb = 0; tb < 2; ++tb) {
      const int m = 2 * pr + tb;
      if (m >= M) continue;
      if (!rp_wait(ev + RP_EV_CLEAR * 2 + tb, target, p.err)) return;
      stamp();
      insert(m);
      rp_arrive(ev + RP_EV_INSERT * 2 + tb);
      stamp();
    }
    for (int tb = 0; tb < 2; ++tb) {
      const int m = 2 * pr + tb;
      if (m >= M) continue;
      if (!rp_wait(ev + RP_EV_INSERT * 2 + tb, target, p.err)) return;
      stamp();
      compact(m);
      rp_arrive(ev + RP_EV_COMPACT * 2 + tb);
      stamp();
    }
    for (int tb = 0; tb < 2; ++tb) {
      const int m = 2 * pr + tb;
      if (m + 2 >= M) continue;
      if (!rp_wait(ev + RP_EV_COMPACT * 2 + tb, target, p.err)) return;
      stamp();
      clear(tb);
      rp_arrive(ev + RP_EV_CLEAR * 2 + tb);
      stamp();
    }
  }
}

struct RpLayout {
  uint32_t slots;
  int log2_slots, groups, tiles_per_tree;
  int64_t n_pad;
  size_t off_events, off_status, off_tables, off_slot_of, total;
};

bool rp_layout(int64_t num_trees, int64_t n_max, RpLayout& L) {
  if (num_trees <= 0 || num_trees >= (1 << 21) || n_max < 0 || n_max >= ((int64_t)1 << 30)) return false;
  uint64_t slots = 1024;
  // TCHGEO_RELABEL_DENSE=1: tables of n_max + 2 slots rounded up (load up to ~0.9, about 0.57 for sampled trees) instead
  // of 2 n_max + 2 (<= 0.5 / ~0.3): half the L2 footprint per tree, so twice the trees in flight, against longer probes
  const char* dense = getenv("TCHGEO_RELABEL_DENSE");
  const uint64_t want = (dense && atoi(dense) != 0) ? (uint64_t)n_max + 2 : 2 * (uint64_t)n_max + 2;
  while (slots < want) slots <<= 1;
  L.slots = (uint32_t)slots;
  L.log2_slots = 0;
  while ((1ull << L.log2_slots) < slots) ++L.log2_slots;
  L.n_pad = (n_max + RP_ITEMS - 1) / RP_ITEMS * RP_ITEMS + RP_ITEMS;
  L.tiles_per_tree = (int)std::max<int64_t>(1, (n_max + RP_TILE - 1) / RP_TILE);
  // groups: as many trees in flight (two per group) as keep tables + slot maps + ids inside the L2 budget
  const size_t per_group = 2 * (slots * 8 + (size_t)L.n_pad * 12);
  const char* e = getenv("TCHGEO_RELABEL_GROUPS");
  int64_t groups = e ? atoi(e) : (int64_t)(((size_t)88 << 20) / std::max<size_t>(per_group, 1));
  groups = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(groups, (num_trees + 1) / 2), RP_MAX_GROUPS));
  L.groups = (int)groups;
  size_t o = 0;
  L.off_events = o; o += rl_align((size_t)RP_MAX_GROUPS * 128);
  L.off_status = o; o += rl_align((size_t)groups * 2 * L.tiles_per_tree * 8);
  L.off_tables = o; o += rl_align((size_t)groups * 2 * slots * 8);
  L.off_slot_of = o; o += rl_align((size_t)groups * 2 * L.n_pad * 4);
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
  p.tiles_per_tree = L.tiles_per_tree;
  p.cap_mask = L.slots - 1; p.hash_shift = 32 - L.log2_slots;
  p.tables = (unsigned long long*)(ws + L.off_tables);
  p.slot_of = (uint32_t*)(ws + L.off_slot_of);
  p.status = (unsigned long long*)(ws + L.off_status);
  p.events = (uint32_t*)(ws + L.off_events);
  p.err = err;
  p.trace = nullptr;
  // arrival counters and look-back words start at zero (epoch 0 = no tree)
  e = cudaMemsetAsync(ws + L.off_events, 0, L.off_tables - L.off_events, stream);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)(p.groups * p.ctas_per_group);
  // debug: TCHGEO_RELABEL_TRACE=<file> dumps per-CTA phase time stamps of the first trees of every call (synchronises)
  static const char* trace_path = getenv("TCHGEO_RELABEL_TRACE");
  const size_t trace_bytes = (size_t)grid * RP_TRACE * 8;
  if (trace_path) {
    e = cudaMalloc(&p.trace, trace_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(p.trace, 0, trace_bytes, stream);
    if (e != cudaSuccess) return e;
  }
  void* args[] = {(void*)&p};
  e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(RP_THREADS), args, 0, stream);
  if (trace_path && e == cudaSuccess) {
    std::vector<unsigned long long> h((size_t)grid * RP_TRACE);
    e = cudaMemcpyAsync(h.data(), p.trace, trace_bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) {
      if (FILE* f = fopen(trace_path, "wb")) {
        const int hdr[4] = {(int)grid, RP_TRACE, p.groups, p.ctas_per_group};
        fwrite(hdr, sizeof(int), 4, f);
        fwrite(h.data(), 8, h.size(), f);
        fclose(f);
      }
    }
    cudaFree(p.trace);
  }
  return e;
}

// =================================================================================================
// Bucketed form (default for ids < 2^32-1): no global hash table at all.
// A tree's ids are split into NB buckets (pairs (id, position), counting sort: count, offsets, scatter), and every bucket
// is resolved by ONE CTA in SHARED MEMORY (shared-memory atomics cost a few cycles, against the ~50 G/s the L2 sustains
// for global atomicCAS, which is what held the persistent form at 7 ms per step): per id the position its map entry
// points at ("winner": the last seed carrying the id, else the first occurrence).  Positions are then walked once in
// order -- flags, block scan, decoupled look-back per tree, node list, local ids of the winners -- and a last pass gives
// later occurrences the local id of their winner.  Two ways to bucket:
//   * direct (the caller states an id bound, as the sampling plan does from the graph): bucket = id mod NB and the
//     bucket's table is a DIRECT-ADDRESS array indexed by id / NB (<= 8192 slots x 4 B) -- one atomicMin per id, no keys,
//     no probing, no capacity to overflow; when bits(bound / NB) + bits(n_max) <= 32 a pair is ONE 32-bit word
//     (slot << pos_bits | position), which halves the scatter's writes and the resolve's reads;
//   * hashed (bound unknown or beyond NB x 8192): bucket = hash(id), open addressing in a 4096 x (key, priority) table.
//     A bucket with more distinct ids than its table holds (hash skew; ~2x headroom over the mean at the worst-case
//     tree size) raises TCHGEO_ERR_CAPACITY rather than a wrong answer.
// Everything is streaming traffic, per id:
//   count 8 B | scatter 8 + 4 (8) B | resolve 4 (8) + 4 B | compact 4 + 8 + ~15 B | lookup 4 + ~2 B
// about 60 B (packed) against the 24 B of the byte model (samples in, nodes + local out), all of it coalesced except the
// 4-byte winner scatter (merged in the L2).  TCHGEO_RELABEL_DIRECT=0 forces the hashed tables,
// TCHGEO_RELABEL_PERSISTENT=1 selects the global-table form.
// =================================================================================================
constexpr int BK_THREADS = 512;
constexpr int BK_ITEMS = 16;
constexpr int BK_TILE = BK_THREADS * BK_ITEMS;    // ids per CTA in the count / scatter kernels
constexpr int BK_RTHREADS = 128;                  // resolve: threads per bucket (many short CTAs in flight)
constexpr int BK_MAX_BUCKETS = 4096;              // shared-memory histogram
constexpr int BK_TABLE = 4096;                    // hashed: slots of a bucket's shared-memory table (32 KB: 7 CTAs per SM)
constexpr int BK_TARGET = 1536;                   // ids per bucket the bucket count is chosen for (worst-case tree)
constexpr int BK_DSLOTS = 8192;                   // direct: slots of a bucket's direct-address table (32 KB)
constexpr int BKC_ITEMS = 8;                      // positions per thread in the compact / lookup kernels
constexpr int BKC_TILE = RL_THREADS * BKC_ITEMS;
constexpr uint32_t BK_NONE = 0xFFFFFFFFu;
enum { BK_HASHED = 0, BK_DIRECT = 1, BK_PACKED = 2 };

struct BkParams {
  const int64_t* samples;
  int64_t stride;
  const int64_t* lens;
  int64_t* nodes;
  int64_t* local;
  int64_t* nodes_len;
  int64_t num_seeds, n_max;
  int32_t num_trees, nb, log2_nb, tiles_per_tree;   // nb buckets per tree; tiles of BK_TILE (count/scatter) positions
  uint32_t id_bound;   // ids are valid below this (hashed: 2^32-1)
  int32_t pos_bits;    // packed pairs: slot << pos_bits | position
  int32_t slots;       // direct: entries of a bucket's table = ceil(id_bound / nb)
  uint32_t* counts;    // [trees, nb] ids per bucket
  uint32_t* offs;      // [trees, nb] exclusive offsets inside the tree's pair region
  uint32_t* tile_hist; // [trees, tiles_per_tree, nb] ids of tile t in bucket j, then (bk_tilescan_kernel) the ids of the
                       // bucket in earlier tiles OF THE SAME SEGMENT of tiles: the scatter needs no atomic cursor
  uint32_t* seg_tot;   // [trees, segs, nb] ids of the bucket in the segment, then (bk_offsets_kernel) in earlier segments
  int32_t segs, tiles_per_seg;   // the tile scan runs per segment of tiles_per_seg tiles (parallelism for few, long trees)
  void* pairs;         // [trees, n_max] uint2 (id, position) or packed uint32, grouped by bucket
  uint32_t* win;       // [trees, n_max] position the map entry of samples[i] points at (BK_NONE: id out of range)
  uint64_t* status;    // [trees, ctiles] look-back words of the compact kernel (zero)
  uint32_t* ticket;    // (zero)
  int32_t ctiles;      // tiles of BKC_TILE positions per tree
  uint32_t* err;
};

__device__ __forceinline__ int64_t bk_len(const BkParams& p, int b) {
  int64_t n = p.lens[b];
  if (n > p.n_max) n = p.n_max;
  return n < 0 ? 0 : n;
}
template <int MODE>
__device__ __forceinline__ uint32_t bk_bucket(uint32_t key, int log2_nb) {
  if (MODE == BK_HASHED) return log2_nb ? (key * 0x9E3779B1u) >> (32 - log2_nb) : 0u;
  return key & ((1u << log2_nb) - 1u);
}

// ---- count: ids of every tile per bucket (a row of tile_hist) ---------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(BK_THREADS) bk_count_kernel(const BkParams p) {
  __shared__ uint32_t s_hist[BK_MAX_BUCKETS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int64_t n = bk_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * BK_TILE;
  if (i0 >= n) return;                              // bk_tilescan_kernel reads the rows of non-empty tiles only
  for (int j = tid; j < p.nb; j += BK_THREADS) s_hist[j] = 0u;
  __syncthreads();
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  int64_t k64[BK_ITEMS];
#pragma unroll
  for (int u = 0; u < BK_ITEMS; ++u) {
    const int64_t i = i0 + u * BK_THREADS + tid;
    k64[u] = i < n ? __ldg(src + i) : 0;
  }
  bool bad = false;
#pragma unroll
  for (int u = 0; u < BK_ITEMS; ++u) {
    if (i0 + u * BK_THREADS + tid >= n) continue;
    if ((uint64_t)k64[u] >= (uint64_t)p.id_bound) bad = true;
    else atomicAdd(&s_hist[bk_bucket<MODE>((uint32_t)k64[u], p.log2_nb)], 1u);
  }
  if (bad) atomicOr(p.err, DEV_ERR_INDEX);
  __syncthreads();
  uint32_t* row = p.tile_hist + ((size_t)b * p.tiles_per_tree + blockIdx.x) * p.nb;
  for (int j = tid; j < p.nb; j += BK_THREADS) row[j] = s_hist[j];
}

// ---- tile scan: per (tree, segment of tiles, bucket) the exclusive prefix over the segment's tiles and its total ---------
__global__ void __launch_bounds__(256) bk_tilescan_kernel(const BkParams p) {
  const int b = blockIdx.z, sg = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j >= p.nb) return;
  const int64_t n = bk_len(p, b);
  const int nt = (int)((n + BK_TILE - 1) / BK_TILE);
  int t = sg * p.tiles_per_seg;
  const int t1 = min(t + p.tiles_per_seg, nt);
  uint32_t* col = p.tile_hist + (size_t)b * p.tiles_per_tree * p.nb + j;
  uint32_t run = 0u;
  constexpr int V = 8;
  for (; t + V <= t1; t += V) {
    uint32_t c[V];
#pragma unroll
    for (int v = 0; v < V; ++v) c[v] = __ldcg(col + (size_t)(t + v) * p.nb);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      col[(size_t)(t + v) * p.nb] = run;
      run += c[v];
    }
  }
  for (; t < t1; ++t) {
    const uint32_t c = __ldcg(col + (size_t)t * p.nb);
    col[(size_t)t * p.nb] = run;
    run += c;
  }
  p.seg_tot[((size_t)b * p.segs + sg) * p.nb + j] = run;
}

// ---- offsets: bucket counts from the segment totals, then their exclusive scan (one CTA per tree; nb <= 4096) ----------
__global__ void __launch_bounds__(256) bk_offsets_kernel(const BkParams p) {
  __shared__ uint32_t s_wtot[8];
  __shared__ uint32_t s_carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0u;
  __syncthreads();
  for (int j0 = 0; j0 < p.nb; j0 += 256) {
    const int j = j0 + tid;
    uint32_t v = 0u;
    if (j < p.nb) {      // segment totals -> ids of the bucket in earlier segments; their sum is the bucket's count
      uint32_t* sg = p.seg_tot + (size_t)b * p.segs * p.nb + j;
      for (int q = 0; q < p.segs; ++q) {
        const uint32_t c = sg[(size_t)q * p.nb];
        sg[(size_t)q * p.nb] = v;
        v += c;
      }
      p.counts[(size_t)b * p.nb + j] = v;
    }
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    uint32_t before = s_carry, total = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += s_wtot[w];
      total += s_wtot[w];
    }
    if (j < p.nb) p.offs[(size_t)b * p.nb + j] = before + incl - v;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
}

// ---- scatter: (id, position) pairs grouped by bucket ---------------------------------------------------------------
// (the form for unpacked pairs: hashed buckets, or direct ones whose slot and position do not fit one word.)  The tile
// scan left in tile_hist the number of ids every bucket received from earlier tiles, so the tile's cursors start at
// their final places and a shared-memory atomicAdd hands out the slots; the order inside a bucket is whatever the
// atomics make it, which the resolve kernel (a minimum per id) does not see.
template <int MODE>
__global__ void __launch_bounds__(BK_THREADS) bk_scatter_kernel(const BkParams p) {
  __shared__ uint32_t s_cur[BK_MAX_BUCKETS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int64_t n = bk_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * BK_TILE;
  if (i0 >= n) return;
  const uint32_t* row = p.tile_hist + ((size_t)b * p.tiles_per_tree + blockIdx.x) * p.nb;
  const uint32_t* seg = p.seg_tot + ((size_t)b * p.segs + blockIdx.x / p.tiles_per_seg) * p.nb;
  for (int j = tid; j < p.nb; j += BK_THREADS)
    s_cur[j] = p.offs[(size_t)b * p.nb + j] + (p.segs > 1 ? seg[j] : 0u) + row[j];
  __syncthreads();
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  uint2* dst = reinterpret_cast<uint2*>(p.pairs) + (size_t)b * p.n_max;
  constexpr int G = 8;                               // ids requested before the first is placed
#pragma unroll 1
  for (int u0 = 0; u0 < BK_ITEMS; u0 += G) {
    int64_t k64[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * BK_THREADS + tid;
      k64[u] = i < n ? __ldg(src + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int64_t i = i0 + (int64_t)(u0 + u) * BK_THREADS + tid;
      if (i >= n) continue;
      if ((uint64_t)k64[u] >= (uint64_t)p.id_bound) {
        p.win[(size_t)b * p.n_max + i] = BK_NONE;    // reported by the count kernel
        continue;
      }
      p.win[(size_t)b * p.n_max + i] = (uint32_t)i;  // "its own winner" until bk_resolve says otherwise (coalesced here,
      const uint32_t key = (uint32_t)k64[u];         // so that the resolve kernel scatters the exceptions only)
      const uint32_t at = atomicAdd(&s_cur[bk_bucket<MODE>(key, p.log2_nb)], 1u);
      dst[at] = make_uint2(key, (uint32_t)i);
    }
  }
}

// ---- scatter, packed pairs: the tile is first grouped by bucket in shared memory, then written out in that order --------
// A warp's store then covers a few runs of consecutive words instead of 32 different sectors.  The kernel is a chain of
// phases (ids, ranks, scan, staging, stores) separated by barriers, so what matters is how many CTAs an SM holds to
// overlap them: 512 threads x 16 ids (a 50 KB staged tile, <= 64 registers) give two, with half as many tiles and tile
// histogram rows as the 8-id form that held three (3.24 -> 3.16 ms for the whole stage); the first version, 1024 threads
// x 61 registers, ran ONE CTA per SM and was no faster than the unstaged scatter.
__global__ void __launch_bounds__(BK_THREADS, 2) bk_scatter_staged_kernel(const BkParams p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  uint32_t* s_start = s_mem;                       // [nb] rank counters, then the bucket's first slot in the staged tile
  uint32_t* s_base = s_mem + p.nb;                 // [nb] bucket's global base minus its first staged slot
  uint32_t* s_stage = s_mem + 2 * p.nb;            // [BK_TILE] packed pairs grouped by bucket
  uint16_t* s_bkt = reinterpret_cast<uint16_t*>(s_stage + BK_TILE);   // [BK_TILE] bucket of every staged pair
  __shared__ uint32_t s_wsum[32];
  __shared__ uint32_t s_total;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = bk_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * BK_TILE;
  if (i0 >= n) return;
  const uint32_t* row = p.tile_hist + ((size_t)b * p.tiles_per_tree + blockIdx.x) * p.nb;
  const uint32_t* seg = p.seg_tot + ((size_t)b * p.segs + blockIdx.x / p.tiles_per_seg) * p.nb;
  const uint32_t* offs = p.offs + (size_t)b * p.nb;
  for (int j = tid; j < p.nb; j += BK_THREADS) {
    s_start[j] = 0u;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(row + j));   // wanted after the scan, two barriers from here
    if (p.segs > 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(seg + j));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(offs + j));
  }
  if (tid < 32) s_wsum[tid] = 0u;
  __syncthreads();
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  uint32_t* win = p.win + (size_t)b * p.n_max;
  const uint32_t bmask = (1u << p.log2_nb) - 1u;
  uint32_t word[BK_ITEMS], rb[BK_ITEMS];           // rb = bucket << 16 | rank in the tile's share of it; ~0 = not staged
  {
    int64_t k64[BK_ITEMS];
#pragma unroll
    for (int u = 0; u < BK_ITEMS; ++u) {
      const int64_t i = i0 + u * BK_THREADS + tid;
      k64[u] = i < n ? __ldg(src + i) : -1;
    }
#pragma unroll
    for (int u = 0; u < BK_ITEMS; ++u) {
      const int64_t i = i0 + u * BK_THREADS + tid;
      rb[u] = BK_NONE; word[u] = 0u;
      if (i >= n) continue;
      if ((uint64_t)k64[u] >= (uint64_t)p.id_bound) {
        win[i] = BK_NONE;                          // reported by the count kernel
        continue;
      }
      win[i] = (uint32_t)i;                        // "its own winner" until bk_resolve says otherwise
      const uint32_t key = (uint32_t)k64[u];
      const uint32_t bkt = key & bmask;
      word[u] = ((key >> p.log2_nb) << p.pos_bits) | (uint32_t)i;
      rb[u] = (bkt << 16) | atomicAdd(&s_start[bkt], 1u);
    }
  }
  __syncthreads();
  // exclusive scan of the tile's bucket counts: thread tid owns buckets [tid * per, tid * per + per)
  const int per = (p.nb + BK_THREADS - 1) / BK_THREADS;
  uint32_t sum = 0u;
  for (int v = 0; v < per; ++v) {
    const int j = tid * per + v;
    if (j < p.nb) sum += s_start[j];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = s_wsum[lane];
    uint32_t wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += up;
    }
    s_wsum[lane] = wi - v;
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  uint32_t excl = s_wsum[warp] + incl - sum;
  for (int v = 0; v < per; ++v) {
    const int j = tid * per + v;
    if (j >= p.nb) break;
    const uint32_t c = s_start[j];
    s_start[j] = excl;
    s_base[j] = __ldcg(offs + j) + (p.segs > 1 ? __ldcg(seg + j) : 0u) + __ldcg(row + j) - excl;
    excl += c;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < BK_ITEMS; ++u) {
    if (rb[u] == BK_NONE) continue;
    const uint32_t bkt = rb[u] >> 16;
    const uint32_t k = s_start[bkt] + (rb[u] & 0xFFFFu);
    s_stage[k] = word[u];
    s_bkt[k] = (uint16_t)bkt;
  }
  __syncthreads();
  uint32_t* dst = reinterpret_cast<uint32_t*>(p.pairs) + (size_t)b * p.n_max;
  const uint32_t total = s_total;
  for (uint32_t k = tid; k < total; k += BK_THREADS) dst[s_base[s_bkt[k]] + k] = s_stage[k];
}

// ---- resolve (direct): one CTA per (tree, bucket); min priority per id in a direct-address shared-memory array -------
// The CTAs are short (a few hundred pairs) and bound by their chain of latencies -- counts, offsets, pairs, atomics,
// stores -- so they are small (128 threads: twice as many in flight per SM) and issue the pair loads before they clear
// the table.
template <bool PACKED>
__global__ void __launch_bounds__(BK_RTHREADS) bk_resolve_direct_kernel(const BkParams p) {
  extern __shared__ __align__(16) uint32_t s_prio[];   // [slots]
  const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const uint32_t cnt = p.counts[(size_t)b * p.nb + j];
  if (cnt == 0) return;
  const size_t base = (size_t)b * p.n_max + p.offs[(size_t)b * p.nb + j];
  const uint32_t* pr1 = reinterpret_cast<const uint32_t*>(p.pairs) + base;
  const uint2* pr2 = reinterpret_cast<const uint2*>(p.pairs) + base;
  const uint32_t S = (uint32_t)p.num_seeds;
  const uint32_t pos_mask = PACKED ? (1u << p.pos_bits) - 1u : 0u;
  uint32_t* win = p.win + (size_t)b * p.n_max;
  constexpr int U = 8;
  uint32_t slot[U], pos[U];
  auto load = [&](uint32_t q0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = q0 + u * BK_RTHREADS + tid;
      pos[u] = BK_NONE; slot[u] = 0u;
      if (q >= cnt) continue;
      if (PACKED) {
        const uint32_t w = __ldcs(pr1 + q);
        slot[u] = w >> p.pos_bits; pos[u] = w & pos_mask;
      } else {
        const uint2 e = __ldcs(pr2 + q);
        slot[u] = e.x >> p.log2_nb; pos[u] = e.y;
      }
    }
  };
  auto vote = [&]() {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pos[u] != BK_NONE) atomicMin(&s_prio[slot[u]], pos[u] < S ? S - 1u - pos[u] : pos[u]);   // smallest priority wins
  };
  auto emit = [&]() {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pos[u] == BK_NONE) continue;
      const uint32_t prio = s_prio[slot[u]];
      const uint32_t w = prio < S ? S - 1u - prio : prio;   // the last seed carrying the id (:26), else its first occurrence
      if (w != pos[u]) win[pos[u]] = w;                     // (the scatter kernel wrote win[i] = i)
    }
  };
  load(0u);
  for (int q = tid; q < p.slots; q += BK_RTHREADS) s_prio[q] = BK_NONE;
  __syncthreads();
  if (cnt <= (uint32_t)BK_RTHREADS * U) {   // the usual case: the bucket's pairs stay in registers between the two phases
    vote();
    __syncthreads();
    emit();
    return;
  }
  vote();
  for (uint32_t q0 = BK_RTHREADS * U; q0 < cnt; q0 += BK_RTHREADS * U) { load(q0); vote(); }
  __syncthreads();
  for (uint32_t q0 = 0; q0 < cnt; q0 += BK_RTHREADS * U) { load(q0); emit(); }   // second read of the bucket: an L2 hit
}

// ---- resolve (hashed): one CTA per (tree, bucket); the bucket's ids in a shared-memory hash table -------------------
__global__ void __launch_bounds__(256) bk_resolve_kernel(const BkParams p) {
  // keys and priorities in two 32-bit arrays: native shared-memory atomicCAS / atomicMin (a 64-bit (key | priority) word
  // needs 64-bit shared atomics, which cost several times more; with those, halving the table to raise the load factor
  // made the kernel 70 % slower)
  extern __shared__ __align__(16) uint32_t s_keys[];   // [BK_TABLE] keys, then [BK_TABLE] priorities (64 KB: dynamic, opt-in)
  uint32_t* s_prio = s_keys + BK_TABLE;
  __shared__ uint32_t s_distinct;
  const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const uint32_t cnt = p.counts[(size_t)b * p.nb + j];
  if (cnt == 0) return;
  for (int q = tid; q < BK_TABLE / 2; q += 256) reinterpret_cast<uint4*>(s_keys)[q] = make_uint4(~0u, ~0u, ~0u, ~0u);
  if (tid == 0) s_distinct = 0u;
  __syncthreads();
  const uint2* pr = reinterpret_cast<const uint2*>(p.pairs) + (size_t)b * p.n_max + p.offs[(size_t)b * p.nb + j];
  const uint32_t S = (uint32_t)p.num_seeds;
  constexpr uint32_t MASK = BK_TABLE - 1, LIMIT = BK_TABLE - BK_TABLE / 8;
  constexpr int U = 8;   // pairs per thread in flight: the kernel is otherwise bound by one DRAM latency per pair
  bool overflow = false;
  for (uint32_t q0 = 0; q0 < cnt && !overflow; q0 += 256 * U) {
    uint2 e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = q0 + u * 256 + tid;
      e[u] = q < cnt ? __ldcs(pr + q) : make_uint2(0u, BK_NONE);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e[u].y == BK_NONE || overflow) continue;
      const uint32_t prio = e[u].y < S ? S - 1u - e[u].y : e[u].y;
      uint32_t h = (e[u].x * 0x85EBCA6Bu) >> (32 - 12);
      const uint32_t step = ((e[u].x * 0xC2B2AE35u) >> 9) | 1u;
      for (int tries = 0;; ++tries) {
        const uint32_t old = atomicCAS(&s_keys[h], BK_NONE, e[u].x);
        if (old == BK_NONE && atomicAdd(&s_distinct, 1u) >= LIMIT) overflow = true;  // keeps probe sequences finite
        if (old == BK_NONE || old == e[u].x) {
          atomicMin(&s_prio[h], prio);   // the smallest priority of the id wins
          break;
        }
        h = (h + step) & MASK;
        if (tries > BK_TABLE) { overflow = true; break; }
      }
    }
  }
  if (__syncthreads_or(overflow)) {
    if (tid == 0) atomicOr(p.err, DEV_ERR_CAPACITY);
    return;
  }
  uint32_t* win = p.win + (size_t)b * p.n_max;
  for (uint32_t q0 = 0; q0 < cnt; q0 += 256 * U) {
    uint2 e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = q0 + u * 256 + tid;
      e[u] = q < cnt ? __ldcs(pr + q) : make_uint2(0u, BK_NONE);   // second read of the bucket: an L2 hit
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (e[u].y == BK_NONE) continue;
      uint32_t h = (e[u].x * 0x85EBCA6Bu) >> (32 - 12);
      const uint32_t step = ((e[u].x * 0xC2B2AE35u) >> 9) | 1u;
      while (s_keys[h] != e[u].x) h = (h + step) & MASK;
      const uint32_t prio = s_prio[h];
      const uint32_t w = prio < S ? S - 1u - prio : prio;   // the last seed carrying the id (:26), else its first occurrence
      if (w != e[u].y) win[e[u].y] = w;                     // (the scatter kernel wrote win[i] = i)
    }
  }
}

// ---- compact: flags, scan in position order (decoupled look-back per tree), node list, local ids of the winners ------
__global__ void __launch_bounds__(RL_THREADS) bk_compact_kernel(const BkParams p) {
  __shared__ uint32_t s_tile;
  __shared__ int64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles in start order and TILE-major (ticket -> tile t of tree ticket % trees): the tiles in flight belong to many
  // trees, so a tile's predecessors in its own tree finished long ago and the look-back finds an inclusive prefix at once
  // (tree-major, all tiles of a tree start together and the late ones sum hundreds of aggregates: 2.2 ms instead of 1.x)
  if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const int t = (int)(s_tile / (uint32_t)p.num_trees), b = (int)(s_tile - (uint32_t)t * (uint32_t)p.num_trees);
  const int64_t n = bk_len(p, b);
  const int64_t i0 = (int64_t)t * BKC_TILE;
  if (i0 >= n && t > 0) return;
  const int64_t* src = p.samples + (int64_t)b * p.stride;
  const uint32_t* win = p.win + (size_t)b * p.n_max;
  int64_t* local = p.local + (int64_t)b * p.stride;
  const int64_t S = p.num_seeds;
  // position of (u, tid) = i0 + u * RL_THREADS + tid: consecutive threads touch consecutive positions in every load and
  // store (a thread owning consecutive positions made every 8-byte store instruction hit 32 different sectors)
  uint32_t w[BKC_ITEMS];
  int64_t id[BKC_ITEMS];    // loaded up front (nearly every position of a sampled tree is a first occurrence): the node
  uint32_t node = 0;        // stores at the end of the kernel then wait for nothing
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + tid;
    w[u] = i < n ? __ldcs(win + i) : BK_NONE;
    id[u] = i < n ? __ldcs(src + i) : 0;
  }
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + tid;
    if (i >= n) continue;
    if (i < S) {                                   // every seed is kept (:25) and maps to the last seed with its id (:26)
      node |= 1u << u;
      st_cs_i64(local + i, w[u] == BK_NONE ? -1 : (int64_t)w[u]);
    } else if (w[u] == BK_NONE) {
      local[i] = -1;
    } else if (w[u] < (uint32_t)S) {
      st_cs_i64(local + i, (int64_t)w[u]);          // a seed carries this id
    } else if (w[u] == (uint32_t)i) {
      node |= 1u << u;                              // first occurrence of an id no seed carries (:36-39)
    }                                               // else: a later occurrence, bk_lookup_kernel
  }
  // exclusive rank of (u, tid) in position order = nodes of rows u' < u  +  nodes of row u in earlier warps / lanes
  __shared__ uint32_t s_cnt[BKC_ITEMS][RL_THREADS / 32];
  uint32_t lane_excl[BKC_ITEMS];
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    const uint32_t m = __ballot_sync(0xffffffffu, (node >> u) & 1u);
    lane_excl[u] = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_cnt[u][warp] = __popc(m);
  }
  __syncthreads();
  uint32_t excl_u[BKC_ITEMS];
  uint32_t total = 0;
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    excl_u[u] = total + lane_excl[u];
#pragma unroll
    for (int k = 0; k < RL_THREADS / 32; ++k) {
      const uint32_t v = s_cnt[u][k];
      if (k < warp) excl_u[u] += v;
      total += v;
    }
  }
  uint64_t* st = p.status + (size_t)b * p.ctiles;
  if (tid == 0) st_relaxed_u64(st + t, (t == 0 ? 2ull << 62 : 1ull << 62) | (uint64_t)total);
  if (warp == 0) {
    int64_t before = 0;
    if (t > 0) {
      int j = t - 1;
      uint32_t spins = 0;
      while (true) {
        const int idx = j - lane;
        const uint64_t v = idx >= 0 ? ld_relaxed_u64(st + idx) : 2ull << 62;
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
        int64_t val = lane <= first_incl ? (int64_t)(v & ((1ull << 62) - 1)) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        before += val;
        if (first_incl < 32) break;
        j -= 32;
      }
      if (lane == 0) st_relaxed_u64(st + t, (2ull << 62) | (uint64_t)(before + total));
    }
    if (lane == 0) s_excl = before;
  }
  __syncthreads();
  const int64_t tile_excl = s_excl;
  if (tid == 0 && i0 + BKC_TILE >= n) p.nodes_len[b] = tile_excl + total;
  int64_t* nodes = p.nodes + (int64_t)b * p.stride;
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    if (!((node >> u) & 1u)) continue;
    const int64_t i = i0 + u * RL_THREADS + tid;
    const uint32_t r = (uint32_t)tile_excl + excl_u[u];
    st_cs_i64(nodes + r, id[u]);
    if (i >= S) local[i] = (int64_t)r;            // (a plain store: bk_lookup_kernel reads it back for the later occurrences)
  }
}

// ---- lookup: later occurrences of a non-seed id take the local id of its first occurrence ----------------------------
__global__ void __launch_bounds__(RL_THREADS) bk_lookup_kernel(const BkParams p) {
  const int b = blockIdx.y;
  const int64_t n = bk_len(p, b);
  const int64_t i0 = (int64_t)blockIdx.x * BKC_TILE;
  if (i0 >= n) return;
  const uint32_t* win = p.win + (size_t)b * p.n_max;
  int64_t* local = p.local + (int64_t)b * p.stride;
  const uint32_t S = (uint32_t)p.num_seeds;
  uint32_t w[BKC_ITEMS];
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    w[u] = (i < n && i >= (int64_t)S) ? __ldcs(win + i) : BK_NONE;
  }
  // local[w] of a first occurrence w was written by bk_compact_kernel and is never written here; all gathers are issued
  // before the first store (reads and writes of the same array: the compiler would otherwise order them pair by pair)
  int64_t r[BKC_ITEMS];
  uint32_t later = 0u;
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u) {
    const int64_t i = i0 + u * RL_THREADS + threadIdx.x;
    r[u] = 0;
    if (w[u] != BK_NONE && w[u] >= S && w[u] != (uint32_t)i) {
      later |= 1u << u;
      r[u] = __ldcg(local + w[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < BKC_ITEMS; ++u)
    if ((later >> u) & 1u) st_cs_i64(local + i0 + u * RL_THREADS + threadIdx.x, r[u]);
}

struct BkLayout {
  int nb, log2_nb, tiles_per_tree, ctiles;
  int mode, pos_bits, slots, segs, tiles_per_seg;
  uint32_t id_bound;
  size_t off_zero, zero_bytes;   // status | ticket: one memset
  size_t off_counts, off_status, off_ticket, off_offs, off_tile_hist, off_seg_tot, off_pairs, off_win, total;
};

inline int bk_bits(uint64_t n) {   // bits needed for values in [0, n)
  int b = 0;
  while (b < 63 && (1ull << b) < n) ++b;
  return b;
}

// id_bound: every valid id is below it (1 .. 2^32-1)
bool bk_layout(int64_t num_trees, int64_t n_max, int64_t id_bound, BkLayout& L) {
  if (num_trees <= 0 || num_trees > 65535 || n_max < 0 || n_max >= ((int64_t)1 << 31)) return false;
  if (id_bound <= 0 || id_bound > 0xFFFFFFFFll) return false;
  int nb = 1, lg = 0;
  while ((int64_t)nb * BK_TARGET < n_max && nb < BK_MAX_BUCKETS) { nb <<= 1; ++lg; }
  const bool hashed_fits = (int64_t)nb * BK_TARGET >= n_max;   // hashed tables: trees up to ~6 M ids
  L.mode = BK_HASHED; L.pos_bits = 0; L.slots = 0; L.id_bound = 0xFFFFFFFFu;
  const char* de = getenv("TCHGEO_RELABEL_DIRECT");
  if (id_bound < 0xFFFFFFFFll && !(de && atoi(de) == 0)) {
    auto slots_for = [&](int l) { return (int64_t)(((uint64_t)id_bound - 1) >> l) + 1; };
    int nbd = nb, lgd = lg;
    while (nbd <= BK_MAX_BUCKETS && slots_for(lgd) > BK_DSLOTS) { nbd <<= 1; ++lgd; }
    if (nbd <= BK_MAX_BUCKETS) {
      L.mode = BK_DIRECT; L.id_bound = (uint32_t)id_bound;
      const int pos_bits = bk_bits((uint64_t)std::max<int64_t>(n_max, 1));
      int nbp = nbd, lgp = lgd;   // a few more (smaller) buckets if that makes a pair fit one word
      while (nbp <= BK_MAX_BUCKETS && nbp <= 4 * nbd && bk_bits((uint64_t)slots_for(lgp)) + pos_bits > 32) { nbp <<= 1; ++lgp; }
      if (nbp <= BK_MAX_BUCKETS && nbp <= 4 * nbd && bk_bits((uint64_t)slots_for(lgp)) + pos_bits <= 32) {
        L.mode = BK_PACKED; L.pos_bits = pos_bits;
        nbd = nbp; lgd = lgp;
      }
      nb = nbd; lg = lgd;
      L.slots = (int)slots_for(lg);
    }
  }
  if (L.mode == BK_HASHED && !hashed_fits) return false;       // the global-table forms
  L.nb = nb; L.log2_nb = lg;
  L.tiles_per_tree = (int)std::max<int64_t>(1, (n_max + BK_TILE - 1) / BK_TILE);
  L.ctiles = (int)std::max<int64_t>(1, (n_max + BKC_TILE - 1) / BKC_TILE);
  if ((int64_t)L.ctiles * num_trees >= ((int64_t)1 << 31)) return false;
  const size_t tb = (size_t)num_trees * nb * 4, nn = (size_t)num_trees * (size_t)std::max<int64_t>(n_max, 1);
  size_t o = 0;
  L.off_zero = o;
  L.off_status = o; o += rl_align((size_t)num_trees * L.ctiles * 8);
  L.off_ticket = o; o += 256;
  L.zero_bytes = o - L.off_zero;
  L.off_counts = o; o += rl_align(tb);
  L.off_offs = o; o += rl_align(tb);
  L.off_tile_hist = o; o += rl_align((size_t)num_trees * L.tiles_per_tree * (size_t)nb * 4);   // depends on the bound
  // segments of the tile scan: only for few, long trees (one tree of 6 M ids: 16 CTAs walking 1500 tiles each took longer
  // than the rest of the stage).  With hundreds of trees one segment is best: 8 segments measured 0.29 ms against 0.22 ms
  // (more rows of the histogram matrix open at once), plus the extra pass over the segment totals.
  const int64_t ctas = (int64_t)((nb + 255) / 256) * num_trees;
  L.segs = ctas >= 1024 ? 1 : (int)std::min<int64_t>((2048 + ctas - 1) / ctas, std::min<int64_t>(64, L.tiles_per_tree));
  L.tiles_per_seg = (L.tiles_per_tree + L.segs - 1) / L.segs;
  L.segs = (L.tiles_per_tree + L.tiles_per_seg - 1) / L.tiles_per_seg;
  L.off_seg_tot = o; o += rl_align((size_t)num_trees * L.segs * (size_t)nb * 4);
  L.off_pairs = o; o += rl_align(nn * 8);
  L.off_win = o; o += rl_align(nn * 4);
  L.total = o + 256;
  return true;
}

tchgeo_status bk_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees, int64_t num_seeds,
                         int64_t n_max, int64_t* nodes, int64_t* local, int64_t* nodes_len, char* ws, const BkLayout& L,
                         uint32_t* err, cudaStream_t stream) {
  BkParams p;
  p.samples = samples; p.stride = stride; p.lens = lens; p.nodes = nodes; p.local = local; p.nodes_len = nodes_len;
  p.num_seeds = num_seeds; p.n_max = n_max; p.num_trees = (int32_t)num_trees;
  p.nb = L.nb; p.log2_nb = L.log2_nb; p.tiles_per_tree = L.tiles_per_tree; p.ctiles = L.ctiles;
  p.id_bound = L.id_bound; p.pos_bits = L.pos_bits; p.slots = L.slots;
  p.counts = (uint32_t*)(ws + L.off_counts); p.tile_hist = (uint32_t*)(ws + L.off_tile_hist);
  p.seg_tot = (uint32_t*)(ws + L.off_seg_tot); p.segs = L.segs; p.tiles_per_seg = L.tiles_per_seg;
  p.offs = (uint32_t*)(ws + L.off_offs); p.pairs = (void*)(ws + L.off_pairs);
  p.win = (uint32_t*)(ws + L.off_win);
  p.status = (uint64_t*)(ws + L.off_status); p.ticket = (uint32_t*)(ws + L.off_ticket);
  p.err = err;
  TCHGEO_CUDA_CHECK(cudaMemsetAsync(ws + L.off_zero, 0, L.zero_bytes, stream));
  const dim3 tiles((unsigned)L.tiles_per_tree, (unsigned)num_trees);
  switch (L.mode) {
    case BK_PACKED: bk_count_kernel<BK_PACKED><<<tiles, BK_THREADS, 0, stream>>>(p); break;
    case BK_DIRECT: bk_count_kernel<BK_DIRECT><<<tiles, BK_THREADS, 0, stream>>>(p); break;
    default: bk_count_kernel<BK_HASHED><<<tiles, BK_THREADS, 0, stream>>>(p); break;
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  bk_tilescan_kernel<<<dim3((unsigned)((L.nb + 255) / 256), (unsigned)L.segs, (unsigned)num_trees), 256, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  bk_offsets_kernel<<<(unsigned)num_trees, 256, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  switch (L.mode) {
    case BK_PACKED: {
      const size_t smem = (size_t)(2 * L.nb + BK_TILE) * 4 + (size_t)BK_TILE * 2;
      static bool configured[64] = {};  // per device; benign race: the attribute is idempotent
      int dev = 0;
      TCHGEO_CUDA_CHECK(cudaGetDevice(&dev));
      if (dev < 0 || dev >= 64 || !configured[dev]) {
        TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(bk_scatter_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (2 * BK_MAX_BUCKETS + BK_TILE) * 4 + BK_TILE * 2));
        if (dev >= 0 && dev < 64) configured[dev] = true;
      }
      bk_scatter_staged_kernel<<<tiles, BK_THREADS, smem, stream>>>(p);
      break;
    }
    case BK_DIRECT: bk_scatter_kernel<BK_DIRECT><<<tiles, BK_THREADS, 0, stream>>>(p); break;
    default: bk_scatter_kernel<BK_HASHED><<<tiles, BK_THREADS, 0, stream>>>(p); break;
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  const dim3 buckets((unsigned)L.nb, (unsigned)num_trees);
  if (L.mode == BK_HASHED) {
    static bool configured[64] = {};  // per device; benign race: the attribute is idempotent
    int dev = 0;
    TCHGEO_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      TCHGEO_CUDA_CHECK(cudaFuncSetAttribute(bk_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_TABLE * 8));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    bk_resolve_kernel<<<buckets, 256, BK_TABLE * 8, stream>>>(p);
  } else if (L.mode == BK_PACKED) {
    bk_resolve_direct_kernel<true><<<buckets, BK_RTHREADS, (size_t)L.slots * 4, stream>>>(p);
  } else {
    bk_resolve_direct_kernel<false><<<buckets, BK_RTHREADS, (size_t)L.slots * 4, stream>>>(p);
  }
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  bk_compact_kernel<<<(unsigned)L.ctiles * (unsigned)num_trees, RL_THREADS, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  bk_lookup_kernel<<<dim3((unsigned)L.ctiles, (unsigned)num_trees), RL_THREADS, 0, stream>>>(p);
  TCHGEO_CUDA_CHECK(cudaGetLastError());
  return TCHGEO_OK;
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

// used by the sampling plan (neighbor_sampling.cu): size and enqueue the stage for one node type.
//   id_bound = 0: any non-negative i64 id -- the wave form with 64-bit keys;
//   0 < id_bound <= 2^32-1: every id is below id_bound (an id outside raises TCHGEO_ERR_INDEX) -- the bucketed form
//   (direct-address shared-memory tables when the bound allows, hashed ones otherwise); TCHGEO_RELABEL_PERSISTENT=1
//   selects the persistent global-table form instead (also the fallback for trees beyond ~6 M ids),
//   TCHGEO_RELABEL_WAVES=1 the wave form.
enum { FORM_WAVES = 0, FORM_PERSISTENT = 1, FORM_BUCKETED = 2 };
static int relabel_form(int64_t num_trees, int64_t n_max, int64_t id_bound, bool prefer_waves = false) {
  const char* w = getenv("TCHGEO_RELABEL_WAVES");
  if (id_bound <= 0 || prefer_waves || (w && atoi(w) != 0)) return FORM_WAVES;
  const char* pe = getenv("TCHGEO_RELABEL_PERSISTENT");
  BkLayout B;
  if (!(pe && atoi(pe) != 0) && bk_layout(num_trees, n_max, id_bound, B)) return FORM_BUCKETED;
  RpLayout P;
  return rp_layout(num_trees, n_max, P) ? FORM_PERSISTENT : FORM_WAVES;
}
bool relabel_is_bucketed(int64_t num_trees, int64_t n_max, int64_t id_bound) {
  return id_bound > 0 && id_bound <= 0xFFFFFFFFll && relabel_form(num_trees, n_max, id_bound) == FORM_BUCKETED;
}
// prefer_waves: the caller has ONE big tree (negative sampling): the wave form, with 8-byte slots when id_bound != 0
size_t relabel_workspace_bytes(int64_t num_trees, int64_t n_max, int64_t id_bound, bool prefer_waves) {
  if (id_bound < 0 || id_bound > 0xFFFFFFFFll) return 0;
  const bool k32 = id_bound != 0;
  RlLayout L;
  if (!rl_layout(num_trees, n_max, k32, L)) return 0;
  const int form = relabel_form(num_trees, n_max, id_bound, prefer_waves);
  if (form == FORM_BUCKETED) {
    BkLayout B;
    bk_layout(num_trees, n_max, id_bound, B);
    return B.total;
  }
  if (form == FORM_PERSISTENT) {
    RpLayout P;
    rp_layout(num_trees, n_max, P);
    return std::max(L.total, P.total);   // (the wave form is its fallback without cooperative launch)
  }
  return L.total;
}
int relabel_launches(int64_t num_trees, int64_t n_max, int64_t id_bound) {
  const int form = relabel_form(num_trees, n_max, id_bound);
  if (form == FORM_BUCKETED) return 7;
  if (form == FORM_PERSISTENT) return 1;
  RlLayout L;
  return rl_layout(num_trees, n_max, id_bound != 0, L) ? 3 * L.num_waves : 0;
}
tchgeo_status relabel_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees,
                              int64_t num_seeds, int64_t n_max, int64_t id_bound, int64_t* nodes, int64_t* local,
                              int64_t* nodes_len, void* workspace, size_t workspace_bytes, uint32_t* err,
                              cudaStream_t stream, bool prefer_waves) {
  TCHGEO_REQUIRE(id_bound >= 0 && id_bound <= 0xFFFFFFFFll, "relabel: id_bound must be 0 (any i64) or at most 2^32-1");
  const bool k32 = id_bound != 0;
  RlLayout L;
  TCHGEO_REQUIRE(rl_layout(num_trees, n_max, k32, L), "relabel: tree too large");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= relabel_workspace_bytes(num_trees, n_max, id_bound, prefer_waves),
                 "relabel: workspace too small (need %zu bytes)",
                 relabel_workspace_bytes(num_trees, n_max, id_bound, prefer_waves));
  const int form = relabel_form(num_trees, n_max, id_bound, prefer_waves);
  if (form == FORM_BUCKETED) {
    BkLayout B;
    bk_layout(num_trees, n_max, id_bound, B);
    return bk_enqueue(samples, stride, lens, num_trees, num_seeds, n_max, nodes, local, nodes_len, (char*)workspace, B, err,
                      stream);
  }
  if (form == FORM_PERSISTENT) {
    RpLayout P;
    rp_layout(num_trees, n_max, P);
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

extern "C" size_t tchgeo_unique_relabel_batched_workspace_bytes(int64_t num_batches, int64_t n_max, int64_t id_bound) {
  return relabel_workspace_bytes(num_batches, n_max, id_bound, false);
}

extern "C" tchgeo_status tchgeo_unique_relabel_batched(const int64_t* samples, int64_t stride, const int64_t* lens,
                                                       int64_t num_batches, int64_t num_seeds, int64_t n_max,
                                                       int64_t id_bound, int64_t* nodes, int64_t* local, int64_t* nodes_len,
                                                       void* workspace, size_t workspace_bytes, int32_t* err_word,
                                                       tchgeo_stream stream_) {
  TCHGEO_REQUIRE(num_batches >= 0 && stride >= 0 && n_max >= 0 && n_max <= stride, "bad relabel geometry");
  TCHGEO_REQUIRE(num_seeds >= 0 && num_seeds <= n_max, "num_seeds out of range");
  if (num_batches == 0) return TCHGEO_OK;
  TCHGEO_REQUIRE(samples && lens && nodes && local && nodes_len && err_word, "NULL pointer");
  return relabel_enqueue(samples, stride, lens, num_batches, num_seeds, n_max, id_bound, nodes, local, nodes_len, workspace,
                         workspace_bytes, (uint32_t*)err_word, (cudaStream_t)stream_, false);
}

extern "C" size_t tchgeo_unique_relabel_workspace_bytes(int64_t n) {
  if (n < 0 || n >= ((int64_t)1 << 30)) return 0;
  return relabel_workspace_bytes(1, n, 0, false);
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
