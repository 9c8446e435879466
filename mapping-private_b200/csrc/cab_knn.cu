// cab_knn.cu -- k-nearest-neighbour mean distances and the statistical outlier filter.
// Replaces the kd-tree loop and the statistics of cloud_algos::StatisticalNoiseRemoval::process
// (cloud_algos/src/noise_removal.cpp:84-136):
//   kdtree_->nearestKSearch(cp, neighborhood_size_, ...) for every point (:86-95)
//   avg[cp] = sum_{ni=1}^{k-1} sqrt(d2[cp][ni]) / (k-1)            (:102-111, the first neighbour is cp itself)
//   mean, stddev of avg (:112-121); keep cp iff |avg - mean| < alpha * stddev (:131)
// The uniform grid answers k-NN queries exactly whenever the k-th neighbour lies within one cell edge of
// the query (every point that close is inside the 3x3x3 cells around it).  Queries with fewer than k
// points that close are retried on a grid with doubled cells until all are resolved; the (d2, index)
// threshold of the k-th neighbour comes from the radix select of cab_topk.cu.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

struct KnnArgs {
  GridView g;
  int p0, p1;
  float r, r2;
  int k;
  const float* thr_d2;   // sorted order: (d2, index) of the k-th neighbour inside the radius, INF if fewer than k + 1
  const int* thr_idx;
  double* avg;           // input order
  unsigned char* done;   // input order
  unsigned int* n_done;
};

__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) knn_mean_kernel(const KnnArgs a) {
  __shared__ ChunkTile tiles[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = a.p0 + blockIdx.x * kWarpsPerBlock + warp;
  if (pid >= a.p1) return;
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
  const int me = g.perm[pc.qi];
  const bool need = pc.active && !a.done[me];
  if (!__any_sync(kFull, need)) return;
  const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
  const float td2 = a.thr_d2[pc.qi];
  const int tidx = a.thr_idx[pc.qi];
  int cnt = 0;
  double sum = 0.0;
  for_each_chunk(g, pc, lane, tile, [&](int n_staged, const float4&, int, bool) {
    for (int m = 0; m < n_staged; ++m) {
      const float d2 = d2_rule(tile->x[m], tile->y[m], tile->z[m], qx, qy, qz);
      if (d2 <= a.r2 && (d2 < td2 || (d2 == td2 && g.perm[tile->idx[m]] <= tidx))) {
        ++cnt;
        sum += (double)__fsqrt_rn(d2);  // std::sqrt(float) at :105; the query itself adds 0 (:104 skips it)
      }
    }
  });
  if (need && cnt == a.k) {
    a.avg[me] = sum / (double)(a.k - 1);
    a.done[me] = 1;
    atomicAdd(a.n_done, 1u);
  }
}

__global__ void knn_init_kernel(const float* __restrict__ xyz, int stride, int n, double* __restrict__ avg,
                                unsigned char* __restrict__ done, unsigned int* __restrict__ n_done) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = xyz + (size_t)i * stride;
  const bool fin = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]);
  avg[i] = __longlong_as_double(0x7ff8000000000000ll);  // NaN until resolved
  done[i] = fin ? 0 : 1;                                // non-finite points never get a value
  if (!fin) atomicAdd(n_done, 1u);
}

// k-NN normals, one round: queries that found their k neighbours inside this round's cell take their result
__global__ void __launch_bounds__(256) knn_normals_collect_kernel(const float4* __restrict__ nrm, const int* __restrict__ kcount,
                                                                  const int* __restrict__ perm, int n_valid, int k,
                                                                  float4* __restrict__ out, unsigned char* __restrict__ done,
                                                                  unsigned int* __restrict__ n_done) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_valid) return;
  const int me = perm[s];
  if (done[me] || kcount[s] != k) return;  // (packets without open queries were skipped: their slots are stale)
  out[me] = nrm[s];
  done[me] = 1;
  atomicAdd(n_done, 1u);
}

}  // namespace

// Edge of the first search grid of a k-NN query: the caller's hint, or an estimate from the cloud's extent.
static int first_knn_cell(cab_ctx* ctx, int k, float cell_hint, const char* who, float* out) {
  float cell = cell_hint;
  if (!(cell > 0.f)) {
    if (int rc = compute_bounds(ctx)) return rc;
    double ext[3] = {0, 0, 0};
    if (ctx->dom_count.empty() || ctx->dom_count[0] == 0) return fail(ctx, CAB_ERR_ARG, "%s: no finite points", who);
    for (int a = 0; a < 3; ++a) {
      ext[a] = std::max(1e-6, (double)ctx->dom_bounds[3 + a] - (double)ctx->dom_bounds[a]);
    }
    const double nv = (double)ctx->dom_count[0];
    // a surface sampled with nv points inside the box has a spacing of about sqrt(largest face / nv) (an
    // under-estimate for volumetric clouds, which only costs doubling rounds); k points need a disc of
    // about sqrt(k / pi) spacings
    const double face = std::max(ext[0] * ext[1], std::max(ext[0] * ext[2], ext[1] * ext[2]));
    const double spacing = std::sqrt(face / nv);
    cell = (float)(spacing * std::sqrt((double)k / M_PI) * 1.5);
  }
  *out = cell;
  return CAB_OK;
}

int run_knn_mean(cab_ctx* ctx, int k, float cell_hint, double* avg_host) {
  if (!ctx->have_cloud) return fail(ctx, CAB_ERR_STATE, "cab_knn_mean_distance: no cloud uploaded");
  if (k < 2) return fail(ctx, CAB_ERR_ARG, "cab_knn_mean_distance: a neighborhood of size %d makes no sense", k);
  if (ctx->n_domains != 1) return fail(ctx, CAB_ERR_ARG, "cab_knn_mean_distance: one cloud at a time");
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (n == 0) return fail(ctx, CAB_ERR_ARG, "cab_knn_mean_distance: %d nearest neighbors requested, but only 0 points in total", k);
  const int saved_rank = ctx->shard_rank, saved_world = ctx->shard_world;
  ctx->shard_rank = 0;
  ctx->shard_world = 1;  // every query is answered here
  struct Restore {
    cab_ctx* c;
    int r, w;
    ~Restore() { c->shard_rank = r; c->shard_world = w; }
  } restore{ctx, saved_rank, saved_world};

  float cell = 0.f;
  if (int rc = first_knn_cell(ctx, k, cell_hint, "cab_knn_mean_distance", &cell)) return rc;
  if ((int64_t)ctx->n < k) return fail(ctx, CAB_ERR_ARG, "cab_knn_mean_distance: %d nearest neighbors requested, but only %lld points in total", k, (long long)ctx->n);

  if (int rc = reserve(ctx, ctx->b_knn_avg, (size_t)n * sizeof(double))) return rc;
  if (int rc = reserve(ctx, ctx->b_knn_done, (size_t)n + 64)) return rc;
  unsigned char* done = (unsigned char*)ctx->b_knn_done.p;
  unsigned int* n_done = (unsigned int*)(done + (((size_t)n + 15) & ~(size_t)15));
  CAB_CUDA(ctx, cudaMemsetAsync(n_done, 0, 4, st));
  knn_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (double*)ctx->b_knn_avg.p, done, n_done);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));

  int rounds = 0;
  for (;; ++rounds) {
    if (rounds > 60) return fail(ctx, CAB_ERR_STATE, "cab_knn_mean_distance: did not converge");
    int rc = build_grid(ctx, cell);
    if (rc == CAB_ERR_OOM) {  // cells too small for the dense table: start coarser
      cell *= 2.f;
      continue;
    }
    if (rc) return rc;
    if (ctx->n_valid < k)
      return fail(ctx, CAB_ERR_ARG, "cab_knn_mean_distance: %d nearest neighbors requested, but only %d points in total", k, ctx->n_valid);
    if (int rc2 = run_thresholds(ctx, cell, k, done)) return rc2;
    KnnArgs a{};
    a.g = grid_view(ctx);
    packet_range(ctx, &a.p0, &a.p1);
    a.r = cell;
    a.r2 = cell * cell;
    a.k = k;
    a.thr_d2 = (const float*)ctx->b_thr_d2.p;
    a.thr_idx = (const int*)ctx->b_thr_idx.p;
    a.avg = (double*)ctx->b_knn_avg.p;
    a.done = done;
    a.n_done = n_done;
    const int np = a.p1 - a.p0;
    if (np > 0) {
      knn_mean_kernel<<<(np + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * kWarp, 0, st>>>(a);
      CAB_LAUNCH_CHECK(ctx);
    }
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, n_done, 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    if (*(const unsigned int*)ctx->h_pin >= (unsigned)n) break;
    cell *= 2.f;
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  if (avg_host) CAB_CUDA(ctx, cudaMemcpyAsync(avg_host, ctx->b_knn_avg.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.knn_ms, ctx->ev[2], ctx->ev[3]));
  ctx->tm.knn_rounds = rounds + 1;
  return CAB_OK;
}

// k-nearest-neighbour normals.  The reference estimates them with nearestKSearch(i, k_) + computePointNormal +
// flipNormalTowardsViewpoint (cloud_tools/src/table_object_detector_passive.cpp:668-714, k_ = 10 at :169;
// cloud_algos/src/cylinder_fit_algo.cpp:138-203): the PCA of cab_normals over the k nearest points (the query
// included, ties by index).  Same rounds as the k-NN mean distance: a query is answered on the first grid whose
// cell holds k points within one edge of it.
int run_normals_knn(cab_ctx* ctx, int k, const float vp[3], float cell_hint, float* out_host) {
  if (!ctx->have_cloud) return fail(ctx, CAB_ERR_STATE, "cab_normals_knn: no cloud uploaded");
  if (k < 3) return fail(ctx, CAB_ERR_ARG, "cab_normals_knn: a plane needs at least 3 points, k = %d", k);
  if (ctx->n_domains != 1) return fail(ctx, CAB_ERR_ARG, "cab_normals_knn: one cloud at a time");
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (n < k) return fail(ctx, CAB_ERR_ARG, "cab_normals_knn: %d nearest neighbors requested, but only %d points in total", k, n);
  const int saved_rank = ctx->shard_rank, saved_world = ctx->shard_world;
  ctx->shard_rank = 0;
  ctx->shard_world = 1;
  struct Restore {
    cab_ctx* c;
    int r, w;
    ~Restore() { c->shard_rank = r; c->shard_world = w; }
  } restore{ctx, saved_rank, saved_world};
  float cell = 0.f;
  if (int rc = first_knn_cell(ctx, k, cell_hint, "cab_normals_knn", &cell)) return rc;

  if (int rc = reserve(ctx, ctx->b_knn_avg, (size_t)n * sizeof(double))) return rc;
  if (int rc = reserve(ctx, ctx->b_knn_done, (size_t)n + 64)) return rc;
  if (int rc = reserve(ctx, ctx->b_out4, (size_t)n * sizeof(float4))) return rc;
  unsigned char* done = (unsigned char*)ctx->b_knn_done.p;
  unsigned int* n_done = (unsigned int*)(done + (((size_t)n + 15) & ~(size_t)15));
  CAB_CUDA(ctx, cudaMemsetAsync(n_done, 0, 4, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_out4.p, 0xff, (size_t)n * sizeof(float4), st));  // NaN until resolved
  knn_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(ctx->xyz_in, ctx->stride, n, (double*)ctx->b_knn_avg.p, done, n_done);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  int rounds = 0;
  float kernels_ms = 0.f;
  for (;; ++rounds) {
    if (rounds > 60) return fail(ctx, CAB_ERR_STATE, "cab_normals_knn: did not converge");
    int rc = build_grid(ctx, cell);
    if (rc == CAB_ERR_OOM) {
      cell *= 2.f;
      continue;
    }
    if (rc) return rc;
    if (ctx->n_valid < k)
      return fail(ctx, CAB_ERR_ARG, "cab_normals_knn: %d nearest neighbors requested, but only %d finite points", k, ctx->n_valid);
    if (int rc2 = run_normals(ctx, cell, k, vp, done)) return rc2;
    kernels_ms += ctx->tm.normals_ms;
    knn_normals_collect_kernel<<<(ctx->n_valid + 255) / 256, 256, 0, st>>>((const float4*)ctx->b_nrm.p, (const int*)ctx->b_kcount.p,
                                                                         (const int*)ctx->b_perm.p, ctx->n_valid, k,
                                                                         (float4*)ctx->b_out4.p, done, n_done);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, n_done, 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    if (*(const unsigned int*)ctx->h_pin >= (unsigned)n) break;
    cell *= 2.f;
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  if (out_host) CAB_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->b_out4.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.knn_ms, ctx->ev[4], ctx->ev[5]));
  ctx->tm.normals_ms = kernels_ms;
  ctx->tm.knn_rounds = rounds + 1;
  ctx->have_normals = false;  // b_nrm holds the last round only; cab_set_normals puts k-NN normals in place for RSD
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_knn_mean_distance(cab_ctx* ctx, int32_t k, float cell_hint, double* avg) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_knn_mean(ctx, k, cell_hint, avg);
}

int cab_normals_knn(cab_ctx* ctx, int32_t k, const float vp[3], float cell_hint, float* nxyz_curv) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_normals_knn(ctx, k, vp, cell_hint, nxyz_curv);
}

int64_t cab_statistical_outliers(cab_ctx* ctx, int32_t k, double alpha, float cell_hint, uint8_t* keep, double* avg,
                                 double* mean_out, double* stddev_out) {
  if (!ctx) return CAB_ERR_ARG;
  if (alpha < 0) return fail(ctx, CAB_ERR_ARG, "cab_statistical_outliers: a STD limit of %g makes no sense", alpha);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  const int64_t n = ctx->n;
  std::vector<double> local;
  double* a = avg;
  if (!a) {
    local.resize((size_t)std::max<int64_t>(n, 1));
    a = local.data();
  }
  if (int rc = run_knn_mean(ctx, k, cell_hint, a)) return rc;
  // statistics in the reference's order (noise_removal.cpp:112-121), fp64, over the finite points
  double sum = 0, sq_sum = 0;
  int64_t m = 0;
  for (int64_t cp = 0; cp < n; ++cp) {
    if (std::isnan(a[cp])) continue;
    sum += a[cp];
    sq_sum += a[cp] * a[cp];
    ++m;
  }
  const double mean = m ? sum / (double)m : 0.0;
  const double variance = m ? sq_sum / (double)m - mean * mean : 0.0;
  const double stddev = std::sqrt(variance);
  int64_t kept = 0;
  for (int64_t cp = 0; cp < n; ++cp) {
    const bool kp = !std::isnan(a[cp]) && std::fabs(a[cp] - mean) < alpha * stddev;  // :131
    if (keep) keep[cp] = kp ? 1 : 0;
    kept += kp ? 1 : 0;
  }
  if (mean_out) *mean_out = mean;
  if (stddev_out) *stddev_out = stddev;
  return kept;
}

}  // extern "C"
