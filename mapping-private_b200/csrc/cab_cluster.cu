// cab_cluster.cu -- Euclidean clustering on the device: the step that cuts the object clusters GRSD runs on
// out of the points above a table.
// Replaces cloud_geometry::nearest::extractEuclideanClusters(points, indices, tolerance, clusters, -1, -1,
// -1, -1, min_pts) [point_cloud_mapping, external] as called at
// cloud_tools/src/table_object_detector_passive.cpp:293,567 and cloud_tools/src/table_object_detector_sr.cpp:370:
// a breadth-first growth through radiusSearch(tolerance) from every unprocessed point in index order, i.e.
// the connected components of the graph "d2 <= tolerance^2" numbered by their smallest index, components
// with fewer than min_pts points dropped.
//
// Here: union-find over the sorted order.  One traversal of the search grid visits every edge once (from
// its higher endpoint); a staged copy of the candidates' parent pointers filters the edges whose endpoints
// are already known to be in one tree, so the lock-free union (CAS hooking of the larger root under the
// smaller, path halving) only runs for the few edges that can still merge two trees.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

// parent pointers are read around L1: a stale "I am a root" would make the CAS below fail forever
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = __ldcg(parent + x);
  while (p != x) {
    const int gp = __ldcg(parent + p);
    if (gp != p) parent[x] = gp;  // path halving; x is not a root and never becomes one again
    x = p;
    p = gp;
  }
  return x;
}

// merges the trees of a and b; parent[x] <= x always holds, so no cycles.  Returns the surviving root.
__device__ __forceinline__ int uf_unite(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return a;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(parent + a, a, b) == a) return b;  // a was still a root: hooked under b
  }
}

struct ClusterArgs {
  GridView g;
  int p0, p1;
  float r, r2;
  int* parent;               // sorted order
  unsigned long long* stats;
};

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) cluster_union_kernel(const ClusterArgs a) {
  __shared__ ChunkTile tiles[kWarpsPerBlock];
  __shared__ int staged_parent[kWarpsPerBlock][kWarp];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  int* spar = staged_parent[warp];
  for (int pid = a.p0 + next_packet(a.stats, lane); pid < a.p1; pid = a.p0 + next_packet(a.stats, lane)) {
    const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
    const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
    int my_root = pc.qi;
    for_each_chunk(g, pc, lane, tile, [&](int, const float4&, int j, bool valid) {
      spar[lane] = valid ? __ldcg(a.parent + j) : -1;
      __syncwarp();
      unsigned mask = pc.active ? chunk_hit_mask(tile, qx, qy, qz, a.r2) : 0u;
      while (mask) {
        const int m = __ffs(mask) - 1;
        mask &= mask - 1;
        const int j2 = tile->idx[m];
        if (j2 >= pc.qi) continue;            // every edge once, from its higher endpoint; skips the query itself
        if (spar[m] == my_root) continue;     // already in my tree (my_root was a root of my tree at some time)
        my_root = uf_unite(a.parent, my_root, j2);
      }
      __syncwarp();
    });
  }
}

__global__ void iota_kernel(int* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// Root of every point, size and smallest input index of every tree.  The roots go to their own array: a root written
// back into parent[s] could be overwritten by the path halving of a thread that walks through s at the same time (with
// an ancestor it read earlier, not the root), and the kernels after this one need the root itself.
// Lanes of a warp that share a root (the usual case: neighbours in the sorted order) combine before the atomics.
__global__ void __launch_bounds__(256) cluster_stats_kernel(int* __restrict__ parent, const int* __restrict__ perm, int n_valid,
                                                            int* __restrict__ root_of, int* __restrict__ size,
                                                            int* __restrict__ min_idx) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = s < n_valid;
  const int root = in ? uf_find(parent, s) : -1 - (int)(threadIdx.x & 31);
  const int idx = in ? perm[s] : 0x7fffffff;
  const unsigned peers = __match_any_sync(kFull, root);
  const int lo = __reduce_min_sync(peers, idx);
  if (in) {
    root_of[s] = root;
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
      atomicAdd(size + root, __popc(peers));
      atomicMin(min_idx + root, lo);
    }
  }
}

// a kept tree marks its smallest input index: the exclusive prefix of the marks numbers the clusters in
// the order the reference's seed loop would have found them
__global__ void __launch_bounds__(256) cluster_seed_kernel(const int* __restrict__ root_of, const int* __restrict__ size,
                                                           const int* __restrict__ min_idx, int n_valid, int min_pts, int max_pts,
                                                           int* __restrict__ seed_flag) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_valid || root_of[s] != s) return;
  const int sz = size[s];
  if (sz >= min_pts && (max_pts <= 0 || sz <= max_pts)) seed_flag[min_idx[s]] = 1;
}

__global__ void __launch_bounds__(256) cluster_label_kernel(const int* __restrict__ root_of, const int* __restrict__ perm,
                                                            const int* __restrict__ size, const int* __restrict__ min_idx,
                                                            const int* __restrict__ seed_rank, int n_valid, int min_pts,
                                                            int max_pts, int* __restrict__ labels) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_valid) return;
  const int root = root_of[s];
  const int sz = size[root];
  labels[perm[s]] = (sz >= min_pts && (max_pts <= 0 || sz <= max_pts)) ? seed_rank[min_idx[root]] : -1;
}

}  // namespace

int64_t run_euclidean_clusters(cab_ctx* ctx, double tolerance, int min_pts, int max_pts, int32_t* labels_host) {
  if (!ctx->have_cloud) return fail(ctx, CAB_ERR_STATE, "cab_euclidean_clusters: no cloud uploaded");
  if (!(tolerance > 0) || !std::isfinite(tolerance)) return fail(ctx, CAB_ERR_ARG, "cab_euclidean_clusters: tolerance must be > 0");
  if (ctx->n_domains != 1) return fail(ctx, CAB_ERR_ARG, "cab_euclidean_clusters: one cloud at a time");
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (n == 0) return 0;
  // connectivity is global: the whole cloud is clustered on this GPU whatever the shard setting
  const int saved_rank = ctx->shard_rank, saved_world = ctx->shard_world;
  ctx->shard_rank = 0;
  ctx->shard_world = 1;
  struct Restore {
    cab_ctx* c;
    int r, w;
    ~Restore() { c->shard_rank = r; c->shard_world = w; }
  } restore{ctx, saved_rank, saved_world};
  const float tol = (float)tolerance;
  if (int rc = build_grid(ctx, tol)) return rc;
  const int nv = ctx->n_valid;

  // parent | root_of | size | min_idx (sorted order), seed_flag | seed_rank (input order, n + 1), labels (input order)
  const size_t n1 = (size_t)n + 1;
  if (int rc = reserve(ctx, ctx->b_cluster, (4 * (size_t)std::max(nv, 1) + 2 * n1 + (size_t)n) * sizeof(int))) return rc;
  if (int rc = reserve(ctx, ctx->b_stats, kStatBytes)) return rc;
  int* parent = (int*)ctx->b_cluster.p;
  int* root_of = parent + std::max(nv, 1);
  int* size = root_of + std::max(nv, 1);
  int* min_idx = size + std::max(nv, 1);
  int* seed_flag = min_idx + std::max(nv, 1);
  int* seed_rank = seed_flag + n1;
  int* labels = seed_rank + n1;
  size_t tmp_scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (const int*)nullptr, (int*)nullptr, (int)n1, st);
  if (int rc = reserve(ctx, ctx->b_cubtmp, tmp_scan + 16)) return rc;
  if (int rc = reserve_pinned(ctx, 64)) return rc;

  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_stats.p, 0, kStatBytes, st));
  CAB_CUDA(ctx, cudaMemsetAsync(size, 0, (size_t)std::max(nv, 1) * sizeof(int), st));
  CAB_CUDA(ctx, cudaMemsetAsync(min_idx, 0x7f, (size_t)std::max(nv, 1) * sizeof(int), st));
  CAB_CUDA(ctx, cudaMemsetAsync(seed_flag, 0, n1 * sizeof(int), st));
  CAB_CUDA(ctx, cudaMemsetAsync(labels, 0xff, (size_t)n * sizeof(int), st));  // non-finite points: -1
  if (nv > 0) {
    iota_kernel<<<(nv + 255) / 256, 256, 0, st>>>(parent, nv);
    CAB_LAUNCH_CHECK(ctx);
    ClusterArgs a{};
    a.g = grid_view(ctx);
    a.p0 = 0;
    a.p1 = ctx->n_packets;
    a.r = tol;
    a.r2 = tol * tol;
    a.parent = parent;
    a.stats = (unsigned long long*)ctx->b_stats.p;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)cluster_union_kernel, kWarpsPerBlock * kWarp, 0);
    const unsigned blocks = (unsigned)std::min<long long>((long long)std::max(per_sm, 1) * ctx->sm_count,
                                                          (ctx->n_packets + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (ctx->n_packets > 0) {
      cluster_union_kernel<<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a);
      CAB_LAUNCH_CHECK(ctx);
    }
    const unsigned gv = (unsigned)((nv + 255) / 256);
    cluster_stats_kernel<<<gv, 256, 0, st>>>(parent, (const int*)ctx->b_perm.p, nv, root_of, size, min_idx);
    CAB_LAUNCH_CHECK(ctx);
    cluster_seed_kernel<<<gv, 256, 0, st>>>(root_of, size, min_idx, nv, min_pts, max_pts, seed_flag);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp_scan, seed_flag, seed_rank, (int)n1, st));
    ctx->tm.kernel_launches += 2;
    cluster_label_kernel<<<gv, 256, 0, st>>>(root_of, (const int*)ctx->b_perm.p, size, min_idx, seed_rank, nv, min_pts, max_pts, labels);
    CAB_LAUNCH_CHECK(ctx);
  } else {
    CAB_CUDA(ctx, cudaMemsetAsync(seed_rank, 0, n1 * sizeof(int), st));
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, seed_rank + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (labels_host) CAB_CUDA(ctx, cudaMemcpyAsync(labels_host, labels, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.cluster_ms, ctx->ev[2], ctx->ev[3]));
  return (int64_t) * (const int*)ctx->h_pin;
}

}  // namespace cab

using namespace cab;

extern "C" {

int64_t cab_euclidean_clusters(cab_ctx* ctx, double tolerance, int32_t min_pts, int32_t max_pts, int32_t* labels) {
  if (!ctx) return CAB_ERR_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  return run_euclidean_clusters(ctx, tolerance, min_pts, max_pts, labels);
}

int64_t cab_cluster_csr(const int32_t* labels, int64_t n, int32_t n_clusters, int32_t* offsets, int32_t* indices) {
  if (!labels || !offsets || n < 0 || n_clusters < 0) return CAB_ERR_ARG;
  for (int32_t c = 0; c <= n_clusters; ++c) offsets[c] = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (labels[i] >= n_clusters) return CAB_ERR_ARG;
    if (labels[i] >= 0) offsets[labels[i] + 1]++;
  }
  for (int32_t c = 0; c < n_clusters; ++c) offsets[c + 1] += offsets[c];
  if (indices) {
    std::vector<int32_t> fill(offsets, offsets + n_clusters);
    for (int64_t i = 0; i < n; ++i)
      if (labels[i] >= 0) indices[fill[labels[i]]++] = (int32_t)i;  // ascending inside every cluster
  }
  return offsets[n_clusters];
}

}  // extern "C"
