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
inline size_t rl_align(size_t x) { return (x + 255) / 256 * 256; }

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

// used by the sampling plan (neighbor_sampling.cu): size and enqueue the stage for one node type
size_t relabel_workspace_bytes(int64_t num_trees, int64_t n_max, bool k32) {
  RlLayout L;
  return rl_layout(num_trees, n_max, k32, L) ? L.total : 0;
}
int relabel_launches(int64_t num_trees, int64_t n_max, bool k32) {
  RlLayout L;
  return rl_layout(num_trees, n_max, k32, L) ? 3 * L.num_waves : 0;
}
tchgeo_status relabel_enqueue(const int64_t* samples, int64_t stride, const int64_t* lens, int64_t num_trees,
                              int64_t num_seeds, int64_t n_max, bool k32, int64_t* nodes, int64_t* local,
                              int64_t* nodes_len, void* workspace, size_t workspace_bytes, uint32_t* err,
                              cudaStream_t stream) {
  RlLayout L;
  TCHGEO_REQUIRE(rl_layout(num_trees, n_max, k32, L), "relabel: tree too large");
  TCHGEO_REQUIRE(workspace && workspace_bytes >= L.total, "relabel: workspace too small (need %zu bytes)", L.total);
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
