// ksg_bundle_order.cuh — `merged`: bundle order = iteration order of the reference's std::unordered_map (merged.cpp:210-231),
// all rehash phases in ONE launch.
//
// libstdc++ keeps one singly linked node list: an insertion into an empty bucket goes to the list FRONT, into a non-empty bucket
// to the front of that bucket's run, and a rehash re-inserts every node in list order by the same two rules.  Hence every phase
// (rehash + the insertions up to the next rehash) orders the nodes by
//     (arrival of the FIRST node of the node's bucket, descending ; own arrival, descending)
// with arrival = position in the old list for rehashed nodes, then insertion time (tools/libstdcxx_order.py,
// tests/test_unordered_map_order.py prove this against the real container).  Distinct buckets have distinct first arrivals, so
// the position of a node in the new list is
//     (number of nodes in buckets whose first arrival is later)  +  (number of nodes of its own bucket that arrived later)
// = an exclusive suffix sum over "bucket size, stored at the bucket's first arrival" plus a rank inside the (short) bucket list:
// a counting sort, no comparison sort and no host round trip.  One thread-block cluster (8 CTAs x 1024 threads, hardware
// cluster barrier between the passes) handles one of the reference's two maps (voxel_map / clear_map); the two clusters of the
// launch run concurrently.  Round 1 drove ~14 phases x (memset + 2 kernels + CUB radix sort) per map from the host: 1.3 ms per
// 640x480 / 2 cm frame; this kernel replaces all of it.
#pragma once
#include <cooperative_groups.h>

#include "ksg_kernels.cuh"

namespace ksg {
namespace cg = cooperative_groups;

static constexpr int kBordCluster = 8;
static constexpr int kBordThreads = 1024;

struct BordBuf {
  const int* phase_start;        // [n_phases] insertion index at which the bucket count changes (rehash schedule of the platform's libstdc++)
  const uint32_t* phase_buckets; // [n_phases] bucket count of the phase
  int n_phases;
  uint32_t bucket_cap;           // entries per map in first / head
  const uint32_t* hash;          // LongIndexHash per bundle, canonical order, voxel_map bundles first
  int *ord_a, *ord_b;            // list order (ping-pong): position -> node
  int *first, *head;             // per bucket: earliest arrival, member list head        [2 * bucket_cap]
  int *next, *size_at, *rank;    // per arrival
  int* cta_tot;                  // [2 * kBordCluster]
};

// exclusive suffix sum of a[0..m) in place: a[t] <- sum of a[t'] for t' > t   (cluster-wide)
__device__ __forceinline__ void bord_suffix_scan(cg::cluster_group& cluster, int* a, int m, int* cta_tot, int gt, int nthreads, int crank) {
  __shared__ int s_warp[32];
  const int per = (m + nthreads - 1) / nthreads;
  const int u0 = min(m, gt * per), u1 = min(m, u0 + per);   // u = m - 1 - t
  int local = 0;
  for (int u = u0; u < u1; ++u) local += __ldcg(&a[m - 1 - u]);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
    s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int block_excl = incl - local + (wid > 0 ? s_warp[wid - 1] : 0);
  if (threadIdx.x == 0) __stcg(&cta_tot[crank], s_warp[31]);
  cluster.sync();
  int run = block_excl;
  for (int c = 0; c < crank; ++c) run += __ldcg(&cta_tot[c]);
  for (int u = u0; u < u1; ++u) { const int v = __ldcg(&a[m - 1 - u]); __stcg(&a[m - 1 - u], run); run += v; }
}

__global__ void __cluster_dims__(kBordCluster, 1, 1) __launch_bounds__(kBordThreads, 1)
k_bundle_order(const Counters* __restrict__ cnt, BordBuf bb, const int* __restrict__ bundle_f, int* __restrict__ bundle_f_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const int map_id = blockIdx.x / kBordCluster;       // 0: voxel_map (merged.cpp:126-134), 1: clear_map (merged.cpp:138-145)
  const int crank = (int)cluster.block_rank();
  const int nthreads = kBordCluster * kBordThreads;
  const int gt = crank * kBordThreads + threadIdx.x;
  const int nb_all = cnt->n_cast;
  if (nb_all <= 0) return;
  int nb_vox = cnt->n_nonclear;
  if (nb_vox < 0) nb_vox = 0;
  if (nb_vox > nb_all) nb_vox = nb_all;
  const int off = map_id == 0 ? 0 : nb_vox;
  const int n = map_id == 0 ? nb_vox : nb_all - nb_vox;
  if (n <= 0) return;                                  // uniform over the cluster
  const uint32_t* hash = bb.hash + off;
  int *cur = bb.ord_a + off, *nxt = bb.ord_b + off;
  int *next = bb.next + off, *size_at = bb.size_at + off, *rank = bb.rank + off;
  int *first = bb.first + (size_t)map_id * bb.bucket_cap, *head = bb.head + (size_t)map_id * bb.bucket_cap;
  int* cta_tot = bb.cta_tot + map_id * kBordCluster;
  int n_old = 0;
  for (int p = 0; p < bb.n_phases && n_old < n; ++p) {
    const uint32_t B = bb.phase_buckets[p];
    const int end = (p + 1 < bb.n_phases) ? bb.phase_start[p + 1] : 0x7fffffff;
    const int m = end < n ? end : n;
    for (uint32_t b = (uint32_t)gt; b < B; b += (uint32_t)nthreads) { __stcg(&first[b], 0x7fffffff); __stcg(&head[b], -1); }
    cluster.sync();
    // arrival t: position in the old list for the nodes that are re-inserted by the rehash, then insertion time
    for (int t = gt; t < m; t += nthreads) {
      const int node = t < n_old ? __ldcg(&cur[t]) : t;
      const uint32_t b = hash[node] % B;
      atomicMin(&first[b], t);
      __stcg(&next[t], atomicExch(&head[b], t));
    }
    cluster.sync();
    for (int t = gt; t < m; t += nthreads) {
      const int node = t < n_old ? __ldcg(&cur[t]) : t;
      const uint32_t b = hash[node] % B;
      int later = 0, total = 0;
      for (int e = __ldcg(&head[b]); e >= 0; e = __ldcg(&next[e])) { ++total; if (e > t) ++later; }
      __stcg(&rank[t], later);
      __stcg(&size_at[t], (__ldcg(&first[b]) == t) ? total : 0);
    }
    cluster.sync();
    bord_suffix_scan(cluster, size_at, m, cta_tot, gt, nthreads, crank);
    cluster.sync();
    for (int t = gt; t < m; t += nthreads) {
      const int node = t < n_old ? __ldcg(&cur[t]) : t;
      const uint32_t b = hash[node] % B;
      __stcg(&nxt[__ldcg(&size_at[__ldcg(&first[b])]) + __ldcg(&rank[t])], node);
    }
    cluster.sync();
    int* sw = cur; cur = nxt; nxt = sw;
    n_old = m;
  }
  for (int pos = gt; pos < n; pos += nthreads) bundle_f_out[off + pos] = bundle_f[off + __ldcg(&cur[pos])];
}

// Record ranges of the bundles: b_base[b] = sum of nsteps[b'] for b' < b, in RANK order (exclusive scan by one cluster).  The records
// of a frame are therefore laid out by (bundle rank, ray step), and a STABLE sort on the voxel bits alone leaves every voxel's
// records in rank order = the reference's per-voxel update order: four radix passes instead of seven.
__global__ void __cluster_dims__(kBordCluster, 1, 1) __launch_bounds__(kBordThreads, 1)
k_bundle_scan(Counters* cnt, int* __restrict__ nsteps, long long* __restrict__ b_base, long long rec_cap, unsigned long long* cta_tot) {
  __shared__ unsigned long long s_warp[32];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int nthreads = kBordCluster * kBordThreads;
  const int gt = crank * kBordThreads + threadIdx.x;
  const int n = cnt->n_cast;
  const int per = (n + nthreads - 1) / nthreads;
  const int i0 = min(n, gt * per), i1 = min(n, i0 + per);
  unsigned long long local = 0;
  for (int i = i0; i < i1; ++i) local += (unsigned long long)nsteps[i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long incl = local;
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned long long w = s_warp[lane];
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
    s_warp[lane] = w;
  }
  __syncthreads();
  const unsigned long long block_excl = incl - local + (wid > 0 ? s_warp[wid - 1] : 0ull);
  if (threadIdx.x == 0) __stcg(&cta_tot[crank], s_warp[31]);
  cluster.sync();
  unsigned long long run = block_excl, total = 0;
  for (int c = 0; c < kBordCluster; ++c) { const unsigned long long t = __ldcg(&cta_tot[c]); if (c < crank) run += t; total += t; }
  const bool fits = total <= (unsigned long long)rec_cap;
  for (int i = i0; i < i1; ++i) {
    b_base[i] = (long long)run;
    run += (unsigned long long)nsteps[i];
    if (!fits) nsteps[i] = 0;
  }
  if (gt == 0) { if (fits) cnt->n_records = total; else { cnt->n_records = 0; set_err(cnt, 4); } }
}

}  // namespace ksg
