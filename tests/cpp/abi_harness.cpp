// C++ caller of the C ABI (include/tchgeo_cuda.h) without Python or torch: the proof that a non-Python host can drive
// the path exactly as the reference's Rust host would through `extern "C"`.
//
//   abi_harness <edges.bin> <num_nodes>          edges.bin = int64 [2, E] row-major (rows = src, cols = dst)
//
// It builds the CSC with tchgeo_coo_to_csx, checks the karate colptr anchors of SURVEY 8(c), wraps the arrays in a
// tchgeo_graph_t, creates a tchgeo_plan_t for B = 3 batches of seeds with both fanouts = the maximum degree (17 for
// karate: the deterministic regime, one exact answer) with the relabel stage on, runs enqueue / collect twice on a stream of its
// own, and compares every output with a host-side restatement of src/algo/neighbor_sampling.rs:162-230 and of the
// insertion-order map of src/algo/negative_sampling.rs:20-47.  Exit code 0 = all equal.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "tchgeo_cuda.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                       \
      return 2;                                                                      \
    }                                                                                \
  } while (0)
#define TG(x)                                                                        \
  do {                                                                               \
    tchgeo_status s_ = (x);                                                          \
    if (s_ != TCHGEO_OK) {                                                           \
      fprintf(stderr, "%s -> %d: %s\n", #x, (int)s_, tchgeo_last_error());           \
      return 3;                                                                      \
    }                                                                                \
  } while (0)

template <typename T>
static T* dmalloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
  return (T*)p;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s edges.bin num_nodes\n", argv[0]);
    return 64;
  }
  if (tchgeo_abi_version() != TCHGEO_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch\n");
    return 4;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 65;
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  const int64_t E = bytes / 16, N = atoll(argv[2]);
  std::vector<int64_t> ei((size_t)2 * E);
  if (fread(ei.data(), 8, (size_t)2 * E, f) != (size_t)2 * E) return 66;
  fclose(f);

  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  int64_t *d_row = dmalloc<int64_t>(E), *d_col = dmalloc<int64_t>(E);
  int64_t *d_ptrs = dmalloc<int64_t>(N + 1), *d_idx = dmalloc<int64_t>(E), *d_perm = dmalloc<int64_t>(E);
  CK(cudaMemcpy(d_row, ei.data(), E * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_col, ei.data() + E, E * 8, cudaMemcpyHostToDevice));
  const size_t ws_csx = tchgeo_coo_to_csx_workspace_bytes(E, N, N);
  void* d_ws_csx = dmalloc<char>(ws_csx);
  TG(tchgeo_coo_to_csx(d_row, d_col, E, N, N, 1, d_ptrs, d_idx, d_perm, d_ws_csx, ws_csx, stream));
  CK(cudaStreamSynchronize(stream));
  std::vector<int64_t> ptrs(N + 1), idx(E), perm(E);
  CK(cudaMemcpy(ptrs.data(), d_ptrs, (N + 1) * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(idx.data(), d_idx, E * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(perm.data(), d_perm, E * 8, cudaMemcpyDeviceToHost));
  // host restatement of storage.rs:103-127: perm = argsort(col*N + row), ptrs = ind2ptr(col[perm]), indices = row[perm]
  std::vector<int64_t> order(E);
  for (int64_t e = 0; e < E; ++e) order[e] = e;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return ei[E + a] * N + ei[a] < ei[E + b] * N + ei[b];
  });
  int bad = 0;
  std::vector<int64_t> want_ptrs(N + 1, 0);
  for (int64_t e = 0; e < E; ++e) want_ptrs[ei[E + e] + 1]++;
  for (int64_t i = 0; i < N; ++i) want_ptrs[i + 1] += want_ptrs[i];
  for (int64_t i = 0; i <= N; ++i) bad += ptrs[i] != want_ptrs[i];
  for (int64_t e = 0; e < E; ++e) bad += (perm[e] != order[e]) + (idx[e] != ei[order[e]]);
  if (N == 34 && E == 156) {  // karate anchors
    const int64_t anchor[6] = {0, 16, 25, 35, 41, 44};
    for (int i = 0; i < 6; ++i) bad += ptrs[i] != anchor[i];
    bad += ptrs[34] != 156;
  }
  if (bad) {
    fprintf(stderr, "to_csc mismatch: %d\n", bad);
    return 10;
  }

  // ---- graph handle + plan handle --------------------------------------------------------------------------------
  tchgeo_graph_t* graph = nullptr;
  const int64_t* p_tab[1] = {d_ptrs};
  const int64_t* i_tab[1] = {d_idx};
  const int64_t nmaj[1] = {N}, nnz[1] = {E};
  TG(tchgeo_graph_create(1, p_tab, nmaj, i_tab, nnz, &graph));
  TG(tchgeo_graph_prepare(graph, TCHGEO_PREPARE_INDEX_REPLICA, stream));
  if (tchgeo_graph_derived_bytes(graph) != (size_t)E * 4) return 11;

  const int64_t B = 3, S = 4, H = 2;
  int64_t maxdeg = 1;
  for (int64_t i = 0; i < N; ++i) maxdeg = std::max(maxdeg, ptrs[i + 1] - ptrs[i]);
  const int64_t fan[2] = {maxdeg, maxdeg};  // >= every degree: the deterministic regime (17 for karate)
  const int64_t seeds_h[B * S] = {0, 1, 4, 5, 33, 33, 2, 33, 9, 11, 12, 9};  // batch 1 and 2 carry duplicated seeds
  const int32_t zero = 0;
  tchgeo_sampling_args a = {};
  a.num_node_types = 1; a.num_rels = 1; a.num_hops = (int32_t)H; a.sampler_kind = TCHGEO_SAMPLER_UNIFORM;
  a.rel_src = &zero; a.rel_dst = &zero; a.graph = graph; a.fanouts = fan;
  a.num_batches = B; a.seeds_per_batch = &S;
  int64_t cap_n = 0, cap_e = 0;
  TG(tchgeo_neighbor_sampling_capacity(&a, &cap_n, &cap_e));
  int64_t* d_in = dmalloc<int64_t>(B * S);
  int64_t *d_samples = dmalloc<int64_t>(B * cap_n), *d_nodes = dmalloc<int64_t>(B * cap_n), *d_local = dmalloc<int64_t>(B * cap_n);
  int64_t *d_rows = dmalloc<int64_t>(B * cap_e), *d_cols = dmalloc<int64_t>(B * cap_e), *d_eidx = dmalloc<int64_t>(B * cap_e);
  const int64_t* in_tab[1] = {d_in};
  int64_t *s_tab[1] = {d_samples}, *r_tab[1] = {d_rows}, *c_tab[1] = {d_cols}, *e_tab[1] = {d_eidx};
  int64_t *n_tab[1] = {d_nodes}, *l_tab[1] = {d_local};
  a.inputs = in_tab; a.samples = s_tab; a.samples_stride = &cap_n;
  a.rows = r_tab; a.cols = c_tab; a.edge_index = e_tab; a.edges_stride = &cap_e;
  a.nodes = n_tab; a.local = l_tab;
  a.stream = stream;
  a.workspace_bytes = tchgeo_neighbor_sampling_workspace_bytes(&a);
  if (a.workspace_bytes == 0) return 12;
  a.workspace = dmalloc<char>(a.workspace_bytes);
  tchgeo_plan_t* plan = nullptr;
  TG(tchgeo_plan_create(&a, &plan));
  if (tchgeo_plan_num_launches(plan) < (int)H + 1) return 13;

  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemcpyAsync(d_in, seeds_h, sizeof(seeds_h), cudaMemcpyHostToDevice, stream));
    TG(tchgeo_plan_enqueue(plan, /*seed=*/1234 + rep, /*batch_base=*/7, stream));
    TG(tchgeo_plan_collect(plan));
    const int64_t *slen, *elen, *lo, *nlen;
    TG(tchgeo_plan_results(plan, &slen, &elen, &lo, &nlen));
    std::vector<int64_t> hs(B * cap_n), hr(B * cap_e), hc(B * cap_e), he(B * cap_e), hn(B * cap_n), hl(B * cap_n);
    CK(cudaMemcpy(hs.data(), d_samples, hs.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), d_rows, hr.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.data(), d_cols, hc.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(he.data(), d_eidx, he.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hn.data(), d_nodes, hn.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl.data(), d_local, hl.size() * 8, cudaMemcpyDeviceToHost));
    for (int64_t b = 0; b < B; ++b) {
      // neighbor_sampling.rs:162-230 with fanout >= degree: every neighbour, in CSC order
      std::vector<int64_t> ws(seeds_h + b * S, seeds_h + (b + 1) * S), wr, wc, we, wlo;
      size_t begin = 0, end = ws.size();
      for (int64_t h = 0; h < H; ++h) {
        wlo.push_back((int64_t)ws.size()); wlo.push_back((int64_t)wc.size()); wlo.push_back((int64_t)ws.size());
        for (size_t i = begin; i < end; ++i)
          for (int64_t q = ptrs[ws[i]]; q < ptrs[ws[i] + 1]; ++q) {
            wr.push_back((int64_t)ws.size());
            ws.push_back(idx[q]);
            wc.push_back((int64_t)i);
            we.push_back(q);
          }
        begin = end;
        end = ws.size();
      }
      bad += slen[b] != (int64_t)ws.size() || elen[b] != (int64_t)wc.size();
      for (int64_t h = 0; h < H * 3; ++h) bad += lo[b * H * 3 + h] != wlo[h];
      for (size_t i = 0; i < ws.size() && !bad; ++i) bad += hs[b * cap_n + i] != ws[i];
      for (size_t i = 0; i < wc.size() && !bad; ++i)
        bad += (hr[b * cap_e + i] != wr[i]) + (hc[b * cap_e + i] != wc[i]) + (he[b * cap_e + i] != we[i]);
      // negative_sampling.rs:20-47: seeds first (duplicates kept, map -> last), then first appearances
      std::vector<int64_t> wn(seeds_h + b * S, seeds_h + (b + 1) * S);
      std::map<int64_t, int64_t> m;
      for (int64_t i = 0; i < S; ++i) m[wn[i]] = i;
      for (size_t i = S; i < ws.size(); ++i)
        if (!m.count(ws[i])) {
          m[ws[i]] = (int64_t)wn.size();
          wn.push_back(ws[i]);
        }
      bad += nlen[b] != (int64_t)wn.size();
      for (size_t i = 0; i < wn.size() && !bad; ++i) bad += hn[b * cap_n + i] != wn[i];
      for (size_t i = 0; i < ws.size() && !bad; ++i) bad += hl[b * cap_n + i] != m[ws[i]];
    }
    if (bad) {
      fprintf(stderr, "sampling / relabel mismatch in repetition %d\n", rep);
      return 20;
    }
    // compact host transport: i32 ids + one u8 edge count per node on the bus, the i64 vectors rebuilt on the host
    // (fan-out > 255 cannot travel this way; karate's 17 and fakedataset's maximum degree can)
    if (maxdeg <= 255) {
      int64_t n_tot = 0, e_tot = 0;
      for (int64_t b = 0; b < B; ++b) { n_tot += slen[b]; e_tot += elen[b]; }
      int64_t *d_nl = dmalloc<int64_t>(B), *d_el = dmalloc<int64_t>(B), *d_noff = dmalloc<int64_t>(B + 1), *d_eoff = dmalloc<int64_t>(B + 1);
      int32_t *d_s32 = dmalloc<int32_t>(n_tot + 1), *d_e32 = dmalloc<int32_t>(e_tot + 1), *d_terr = dmalloc<int32_t>(1);
      uint8_t* d_cnt = dmalloc<uint8_t>(n_tot + 1);
      CK(cudaMemcpyAsync(d_nl, slen, B * 8, cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(d_el, elen, B * 8, cudaMemcpyHostToDevice, stream));
      CK(cudaMemsetAsync(d_terr, 0, 4, stream));
      TG(tchgeo_pack_transport(d_samples, cap_n, d_cols, d_eidx, cap_e, d_nl, d_el, B, cap_n, cap_e, d_s32, d_e32, d_cnt, n_tot,
                               d_noff, d_eoff, d_terr, stream));
      std::vector<int32_t> h32(n_tot + 1), he32(e_tot + 1);
      std::vector<uint8_t> hcnt(n_tot + 1);
      std::vector<int64_t> noff(B + 1), eoff(B + 1);
      int32_t terr = -1;
      CK(cudaMemcpyAsync(h32.data(), d_s32, n_tot * 4, cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(he32.data(), d_e32, e_tot * 4, cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(hcnt.data(), d_cnt, n_tot, cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(noff.data(), d_noff, (B + 1) * 8, cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(eoff.data(), d_eoff, (B + 1) * 8, cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(&terr, d_terr, 4, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      if (tchgeo_status_from_error_word((uint32_t)terr) != TCHGEO_OK || noff[B] != n_tot || eoff[B] != e_tot) return 22;
      std::vector<int64_t> us(n_tot + 1), uc(e_tot + 1), ue(e_tot + 1);
      TG(tchgeo_host_unpack_transport(h32.data(), he32.data(), hcnt.data(), noff.data(), eoff.data(), B, us.data(), uc.data(),
                                      ue.data(), 2));
      for (int64_t b = 0; b < B; ++b) {
        for (int64_t i = 0; i < slen[b]; ++i) bad += us[noff[b] + i] != hs[b * cap_n + i];
        for (int64_t i = 0; i < elen[b]; ++i)
          bad += (uc[eoff[b] + i] != hc[b * cap_e + i]) + (ue[eoff[b] + i] != he[b * cap_e + i]);
      }
      if (bad) {
        fprintf(stderr, "compact transport mismatch in repetition %d\n", rep);
        return 23;
      }
      cudaFree(d_nl); cudaFree(d_el); cudaFree(d_noff); cudaFree(d_eoff); cudaFree(d_s32); cudaFree(d_e32); cudaFree(d_terr);
      cudaFree(d_cnt);
    }
  }
  // error path: an out-of-range seed is reported by collect, and the plan stays usable
  const int64_t bad_seeds[B * S] = {0, 1, 4, 5, 33, 33, 2, 33, 9, 11, 12, N + 5};
  CK(cudaMemcpyAsync(d_in, bad_seeds, sizeof(bad_seeds), cudaMemcpyHostToDevice, stream));
  TG(tchgeo_plan_enqueue(plan, 1, 0, stream));
  if (tchgeo_plan_collect(plan) != TCHGEO_ERR_INDEX) return 21;
  tchgeo_plan_destroy(plan);
  tchgeo_graph_destroy(graph);
  printf("abi_harness ok: to_csc, graph handle, plan handle (enqueue/collect x2), 2-hop full-neighbourhood sampling, "
         "relabel and the compact host transport of %lld batches match the host restatement\n", (long long)B);
  return 0;
}
