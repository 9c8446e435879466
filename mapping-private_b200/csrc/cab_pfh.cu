// cab_pfh.cu -- Point Feature Histograms ("star" pair features + the FPFH weighted average).
// Replaces the hot loops of cloud_algos::PointFeatureHistogram::process
// (cloud_algos/src/pfh.cpp:181-350; pair features: cloud_algos/include/cloud_algos/pfh.h:102-238):
//   radius search (<= max_nn nearest, self first and skipped)                         pfh.cpp:181-193
//   alpha, beta, gamma [, delta] of every (point, neighbour) pair in fp64, one 1-D histogram of `quantum`
//   bins per feature, increments of 100 / k                                           pfh.cpp:205-288
//   1/d2-weighted average of the neighbours' histograms (what makes them FPFHs)       pfh.cpp:303-333
//   optional bin-to-bin differences                                                   pfh.cpp:337-350
// Same traversal as the RSD kernel: hit mask per 32-candidate chunk, then only the hits are visited and the
// candidate's position and normal come from the lane that staged it.  The histogram increments are all the
// same value, so a bin is reproduced exactly from its integer count by repeating the reference's
// `float += double` (the order of the neighbours does not matter); the weighted average adds different
// values per neighbour, there the summation order differs from the reference's (ascending distance).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

constexpr int kPfhMaxBins = 64;

struct PfhArgs {
  GridView g;
  int p0, p1;
  float r, r2;
  const float4* nrm;     // sorted order
  const float* thr_d2;   // optional max_nn thresholds
  const int* thr_idx;
  int quantum, nfeat, nbins, stride;  // stride: floats per histogram row (nbins rounded up to a multiple of 4)
  int flags;
  double max_dist;
  float* spfh;           // sorted order, n x stride
  float* out;            // sorted order, n x stride (average pass)
  int* kcount;           // sorted order: neighbours kept (self included)
};

__device__ __forceinline__ double dot3(float a0, float a1, float a2, double b0, double b1, double b2) {
  // float * double products summed left to right, as written at pfh.h:128-130
  return __dadd_rn(__dadd_rn(__dmul_rn((double)a0, b0), __dmul_rn((double)a1, b1)), __dmul_rn((double)a2, b2));
}

// pfh.h:102-238 for source point s (the query) and target t (the neighbour).  false: invalid pair.
__device__ __forceinline__ bool pair_features(float sx, float sy, float sz, float snx, float sny, float snz, float tx, float ty,
                                              float tz, float tnx, float tny, float tnz, float d2, double max_dist,
                                              bool check_flip, bool abs_angles, double f[4]) {
  double d0 = (double)__fsub_rn(tx, sx), d1 = (double)__fsub_rn(ty, sy), d2v = (double)__fsub_rn(tz, sz);
  double delta = (double)__fsqrt_rn(d2);
  if (delta <= 0) {
    const double dsq = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2v, d2v));
    if (dsq == 0) return false;
    delta = sqrt(dsq);
  }
  const double angle2 = -dot3(tnx, tny, tnz, d0, d1, d2v) / delta;
  bool flip = !check_flip;
  double gamma = 0;
  if (check_flip) {
    gamma = dot3(snx, sny, snz, d0, d1, d2v) / delta;
    if (acos(gamma) > acos(angle2)) flip = true;
  }
  double u0 = snx, u1 = sny, u2 = snz, n0 = tnx, n1 = tny, n2 = tnz;  // u: source normal, n: target normal
  if (flip) {
    u0 = tnx; u1 = tny; u2 = tnz;
    n0 = snx; n1 = sny; n2 = snz;
    d0 = -d0; d1 = -d1; d2v = -d2v;
    gamma = angle2;
  }
  if (abs_angles) gamma = fabs(gamma);
  const double t0 = __dsub_rn(__dmul_rn(d1, u2), __dmul_rn(d2v, u1));
  const double t1 = __dsub_rn(__dmul_rn(d2v, u0), __dmul_rn(d0, u2));
  const double t2 = __dsub_rn(__dmul_rn(d0, u1), __dmul_rn(d1, u0));
  const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(t0, t0), __dmul_rn(t1, t1)), __dmul_rn(t2, t2)));
  if (nrm == 0) return false;
  const double v0 = t0 / nrm, v1 = t1 / nrm, v2 = t2 / nrm;
  const double w0 = __dsub_rn(__dmul_rn(u1, v2), __dmul_rn(u2, v1));
  const double w1 = __dsub_rn(__dmul_rn(u2, v0), __dmul_rn(u0, v2));
  const double w2 = __dsub_rn(__dmul_rn(u0, v1), __dmul_rn(u1, v0));
  double beta = __dadd_rn(__dadd_rn(__dmul_rn(v0, n0), __dmul_rn(v1, n1)), __dmul_rn(v2, n2));
  if (abs_angles) beta = fabs(beta);
  const double wy = __dadd_rn(__dadd_rn(__dmul_rn(w0, n0), __dmul_rn(w1, n1)), __dmul_rn(w2, n2));
  const double ux = __dadd_rn(__dadd_rn(__dmul_rn(u0, n0), __dmul_rn(u1, n1)), __dmul_rn(u2, n2));
  double alpha = abs_angles ? atan2(fabs(wy), fabs(ux)) : atan2(wy, ux);
  delta = delta / max_dist;
  if (abs_angles) {
    alpha = alpha / (M_PI / 2);
  } else {
    alpha = __dadd_rn(alpha, M_PI) / (2.0 * M_PI);
    beta = __dadd_rn(beta, 1.0) / 2.0;
    gamma = __dadd_rn(gamma, 1.0) / 2.0;
  }
  f[0] = alpha;
  f[1] = beta;
  f[2] = gamma;
  f[3] = delta;
  return true;
}

// Visits the neighbours of every query of a packet: body(m, d2, cx, cy, cz, nx, ny, nz) runs on the lanes whose
// next pending hit is staged candidate m; returns the kept-neighbour count (self included).
template <bool kUseThr, class Body>
__device__ __forceinline__ int visit_neighbours(const PfhArgs& a, const PacketCtx& pc, int lane, ChunkTile* tile, Body&& body) {
  const GridView& g = a.g;
  const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
  float td2 = INFINITY;
  int tidx = INT_MAX;
  if (kUseThr) {
    td2 = a.thr_d2[pc.qi];
    tidx = a.thr_idx[pc.qi];
  }
  int k = 0;
  for_each_chunk(g, pc, lane, tile, [&](int, const float4& c, int j, bool valid) {
    const float4 cn = valid ? a.nrm[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned mask = chunk_hit_mask(tile, qx, qy, qz, a.r2);
    int iters = __reduce_max_sync(kFull, __popc(mask));
#pragma unroll 1
    for (; iters > 0; --iters) {
      const int m = 31 - __clz(mask);
      const float cx = __shfl_sync(kFull, c.x, m), cy = __shfl_sync(kFull, c.y, m), cz = __shfl_sync(kFull, c.z, m);
      const float nx = __shfl_sync(kFull, cn.x, m), ny = __shfl_sync(kFull, cn.y, m), nz = __shfl_sync(kFull, cn.z, m);
      if (mask != 0) {
        mask ^= 1u << m;
        const float d2 = d2_rule(cx, cy, cz, qx, qy, qz);
        bool in = true;
        if (kUseThr) in = d2 < td2 || (d2 == td2 && g.perm[tile->idx[m]] <= tidx);
        if (in) {
          ++k;
          if (tile->idx[m] != pc.qi) body(m, d2, cx, cy, cz, nx, ny, nz);  // the query itself is skipped (pfh.cpp:217)
        }
      }
    }
  });
  return k;
}

template <bool kUseThr>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) spfh_kernel(const PfhArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChunkTile* tiles = reinterpret_cast<ChunkTile*>(smem_raw);
  int* counts = reinterpret_cast<int*>(tiles + kWarpsPerBlock);  // [W][nbins][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = a.p0 + blockIdx.x * kWarpsPerBlock + warp;
  if (pid >= a.p1) return;
  ChunkTile* tile = &tiles[warp];
  int* my = counts + (size_t)warp * a.nbins * kWarp + lane;  // bin b at my[b * 32]
  const PacketCtx pc = load_packet(a.g, pid, lane, a.r, tile);
  const float4 nq = a.nrm[pc.qi];
  for (int b = 0; b < a.nbins; ++b) my[b * kWarp] = 0;
  int ninv = 0;
  const bool check_flip = a.flags & CAB_PFH_CHECK_FLIP, abs_angles = a.flags & CAB_PFH_ABS_ANGLES;
  const int k = visit_neighbours<kUseThr>(a, pc, lane, tile, [&](int, float d2, float cx, float cy, float cz, float nx, float ny, float nz) {
    double f[4];
    if (pair_features(pc.q.x, pc.q.y, pc.q.z, nq.x, nq.y, nq.z, cx, cy, cz, nx, ny, nz, d2, a.max_dist, check_flip, abs_angles, f)) {
      for (int ft = 0; ft < a.nfeat; ++ft) {  // pfh.cpp:224-228: max(0, min(quantum - 1, (int) floor(quantum * feature)))
        const int fi = max(0, min(a.quantum - 1, __double2int_rd(__dmul_rn((double)a.quantum, f[ft]))));
        my[(ft * a.quantum + fi) * kWarp]++;
      }
    } else {
      ++ninv;
    }
  });
  if (pc.active) {
    // every valid pair added npsqr = 100 / k to one bin per feature, every invalid one npsqr / quantum to all bins
    // (pfh.cpp:212,267-286); histograms are floats, the increment a double
    const double npsqr = 100.0 / (double)k, spread = npsqr / (double)a.quantum;
    float* row = a.spfh + (size_t)pc.qi * a.stride;
    for (int b = 0; b < a.nbins; ++b) {
      float h = 0.f;
      for (int c = my[b * kWarp]; c > 0; --c) h = (float)((double)h + npsqr);
      for (int c = ninv; c > 0; --c) h = (float)((double)h + spread);
      row[b] = h;
    }
    for (int b = a.nbins; b < a.stride; ++b) row[b] = 0.f;
    a.kcount[pc.qi] = k;
  }
}

// FPFH step (pfh.cpp:303-333): out[cp][b] = sum_ni spfh[ni][b] * (1 / d2) / sum_ni (1 / d2), floats accumulated with
// double increments like the reference's `float += float * double`.
template <bool kUseThr>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) pfh_average_kernel(const PfhArgs a) {
  __shared__ ChunkTile tiles[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = a.p0 + blockIdx.x * kWarpsPerBlock + warp;
  if (pid >= a.p1) return;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(a.g, pid, lane, a.r, tile);
  float acc[kPfhMaxBins];
#pragma unroll
  for (int b = 0; b < kPfhMaxBins; ++b) acc[b] = 0.f;
  double sum_weight = 0.0;
  const int groups = a.stride >> 2;
  visit_neighbours<kUseThr>(a, pc, lane, tile, [&](int m, float d2, float, float, float, float, float, float) {
    const double weight = 1.0 / (double)d2;
    sum_weight += weight;
    const float4* hn = reinterpret_cast<const float4*>(a.spfh + (size_t)tile->idx[m] * a.stride);
#pragma unroll
    for (int g4 = 0; g4 < kPfhMaxBins / 4; ++g4)
      if (g4 < groups) {
        const float4 h = hn[g4];
        acc[4 * g4 + 0] = (float)((double)acc[4 * g4 + 0] + (double)h.x * weight);
        acc[4 * g4 + 1] = (float)((double)acc[4 * g4 + 1] + (double)h.y * weight);
        acc[4 * g4 + 2] = (float)((double)acc[4 * g4 + 2] + (double)h.z * weight);
        acc[4 * g4 + 3] = (float)((double)acc[4 * g4 + 3] + (double)h.w * weight);
      }
  });
  if (pc.active) {
    float* row = a.out + (size_t)pc.qi * a.stride;
#pragma unroll
    for (int b = 0; b < kPfhMaxBins; ++b)
      if (b < a.stride) row[b] = (float)((double)acc[b] / sum_weight);  // 0 / 0 = NaN for a point without neighbours (:330)
  }
}

// differences (pfh.cpp:337-350) and the way back to input order, point-major rows of nbins floats
__global__ void __launch_bounds__(256) pfh_finish_kernel(const float* __restrict__ rows, int stride, int nbins, int quantum,
                                                         int nfeat, bool differential, const int* __restrict__ perm, int n_valid,
                                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  const int dst = perm[i];
  if (dst < 0) return;
  const float* r = rows + (size_t)i * stride;
  float* o = out + (size_t)dst * nbins;
  for (int ft = 0; ft < nfeat; ++ft) {
    float prev = 0.f;
    for (int b = 0; b < quantum; ++b) {
      const float v = r[ft * quantum + b];
      o[ft * quantum + b] = (differential && b > 0) ? v - prev : v;
      prev = v;
    }
  }
}

}  // namespace

int run_pfh(cab_ctx* ctx, double radius, int max_nn, int quantum, int flags, float* out_host) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_pfh: build the grid first");
  if (!ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "cab_pfh: missing normals");
  if (ctx->shard_world > 1) return fail(ctx, CAB_ERR_STATE, "cab_pfh: the averaging pass reads every neighbour's histogram; not available on a shard");
  const float rf = (float)radius;
  if (!(rf > 0.f) || rf > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_pfh: radius %g exceeds the grid cell %g", radius, (double)ctx->cell);
  const int nfeat = (flags & CAB_PFH_USE_DIST) ? 4 : 3;
  const int nbins = quantum * nfeat;
  if (quantum < 1 || nbins > kPfhMaxBins) return fail(ctx, CAB_ERR_ARG, "cab_pfh: quantum * features must be in [1, %d]", kPfhMaxBins);
  const int n = (int)ctx->n;
  const int stride = (nbins + 3) & ~3;
  cudaStream_t st = ctx->stream;
  if (int rc = reserve(ctx, ctx->b_pfh[0], (size_t)std::max(n, 1) * stride * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_pfh[1], (size_t)std::max(n, 1) * stride * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_pfh[2], (size_t)std::max(n, 1) * nbins * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_kcount, (size_t)std::max(n, 1) * sizeof(int))) return rc;
  const bool use_thr = max_nn > 0;
  if (use_thr)
    if (int rc = run_thresholds(ctx, rf, max_nn)) return rc;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  PfhArgs a{};
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  a.r = rf;
  a.r2 = rf * rf;
  a.nrm = (const float4*)ctx->b_nrm.p;
  a.thr_d2 = use_thr ? (const float*)ctx->b_thr_d2.p : nullptr;
  a.thr_idx = use_thr ? (const int*)ctx->b_thr_idx.p : nullptr;
  a.quantum = quantum;
  a.nfeat = nfeat;
  a.nbins = nbins;
  a.stride = stride;
  a.flags = flags;
  a.max_dist = 2 * radius;  // pfh.cpp:247
  a.spfh = (float*)ctx->b_pfh[0].p;
  a.out = (float*)ctx->b_pfh[1].p;
  a.kcount = (int*)ctx->b_kcount.p;
  const int np = a.p1 - a.p0;
  const float* result = a.spfh;
  if (np > 0) {
    const unsigned blocks = (np + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const size_t smem = (size_t)kWarpsPerBlock * sizeof(ChunkTile) + (size_t)kWarpsPerBlock * nbins * kWarp * sizeof(int);
    if (use_thr) {
      CAB_CUDA(ctx, cudaFuncSetAttribute(spfh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      spfh_kernel<true><<<blocks, kWarpsPerBlock * kWarp, smem, st>>>(a);
    } else {
      CAB_CUDA(ctx, cudaFuncSetAttribute(spfh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      spfh_kernel<false><<<blocks, kWarpsPerBlock * kWarp, smem, st>>>(a);
    }
    CAB_LAUNCH_CHECK(ctx);
    if (flags & CAB_PFH_AVERAGE) {
      if (use_thr) pfh_average_kernel<true><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a);
      else pfh_average_kernel<false><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a);
      CAB_LAUNCH_CHECK(ctx);
      result = a.out;
    }
  }
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_pfh[2].p, 0, (size_t)std::max(n, 1) * nbins * sizeof(float), st));
  if (ctx->n_valid > 0) {
    pfh_finish_kernel<<<(ctx->n_valid + 255) / 256, 256, 0, st>>>(result, stride, nbins, quantum, nfeat, (flags & CAB_PFH_DIFFERENTIAL) != 0,
                                                                 (const int*)ctx->b_perm.p, ctx->n_valid, (float*)ctx->b_pfh[2].p);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  if (out_host && n > 0)
    CAB_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->b_pfh[2].p, (size_t)n * nbins * sizeof(float), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.pfh_ms, ctx->ev[2], ctx->ev[3]));
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_pfh(cab_ctx* ctx, double radius, int32_t max_nn, int32_t quantum, int32_t flags, float* hist) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_pfh(ctx, radius, max_nn, quantum, flags, hist);
}

}  // extern "C"
