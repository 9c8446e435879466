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
#include <cub/device/device_scan.cuh>

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

// ---- combine_ = true: one n-D histogram of quantum ^ features bins per point (pfh.cpp:47-57, 239-258, 279-281) --------
// The bins no longer fit lane-private shared memory (9^3 = 729, 5^4 = 625 ...): a lane counts its query's pairs in the
// query's own row of a global int array (no other lane touches that row), a second kernel turns counts into the
// reference's floats, and the FPFH average runs warp per query over materialised neighbour lists so that the neighbours'
// long rows are read coalesced.
struct PfhCombineArgs {
  int* cnt;              // sorted order, n x nbins pair counts
  int* ninv;             // sorted order: invalid pairs per query
  const int* list_off;   // sorted order: first list entry of the query (exclusive scan of kcount)
  int2* list;            // (sorted index, d2 bits) of every kept neighbour but the query itself
  int slot[4];           // digit of feature f (alpha, beta, gamma, delta) in the bin index (:113-121)
};

template <bool kUseThr, bool kLists>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) pfh_combine_kernel(const PfhArgs a, const PfhCombineArgs ca) {
  __shared__ ChunkTile tiles[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pid = a.p0 + blockIdx.x * kWarpsPerBlock + warp;
  if (pid >= a.p1) return;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(a.g, pid, lane, a.r, tile);
  const float4 nq = a.nrm[pc.qi];
  const bool check_flip = a.flags & CAB_PFH_CHECK_FLIP, abs_angles = a.flags & CAB_PFH_ABS_ANGLES;
  int* row = ca.cnt + (size_t)pc.qi * a.nbins;
  int ninv = 0, filled = 0;
  const int base = kLists ? ca.list_off[pc.qi] : 0;
  const int k = visit_neighbours<kUseThr>(a, pc, lane, tile, [&](int m, float d2, float cx, float cy, float cz, float nx, float ny, float nz) {
    if (!pc.active) return;
    if (kLists) {
      ca.list[base + filled] = make_int2(tile->idx[m], __float_as_int(d2));
      ++filled;
      return;
    }
    double f[4];
    if (pair_features(pc.q.x, pc.q.y, pc.q.z, nq.x, nq.y, nq.z, cx, cy, cz, nx, ny, nz, d2, a.max_dist, check_flip, abs_angles, f)) {
      int index = 0, fi[4] = {0, 0, 0, 0};
      for (int ft = 0; ft < a.nfeat; ++ft)
        fi[ca.slot[ft]] = max(0, min(a.quantum - 1, __double2int_rd(__dmul_rn((double)a.quantum, f[ft]))));
      int power = 1;
      for (int d = 0; d < a.nfeat; ++d) {
        index += power * fi[d];
        power *= a.quantum;
      }
      row[index] += 1;
    } else {
      ++ninv;
    }
  });
  if (pc.active && !kLists) {
    ca.ninv[pc.qi] = ninv;
    a.kcount[pc.qi] = k;
  }
}

// counts -> the reference's floats: every valid pair added npsqr = 100 / k to its bin, every invalid one npsqr / nr_bins
// to all bins (pfh.cpp:212, 251-258, 279-281); histograms are floats, the increment a double
__global__ void __launch_bounds__(256) pfh_combine_fill_kernel(const int* __restrict__ cnt, const int* __restrict__ ninv,
                                                               const int* __restrict__ kcount, int n_valid, int nbins, int stride,
                                                               float* __restrict__ spfh) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_valid * stride) return;
  const int i = (int)(t / stride), b = (int)(t % stride);
  float h = 0.f;
  if (b < nbins) {
    const double npsqr = 100.0 / (double)kcount[i], spread = npsqr / (double)nbins;
    for (int c = cnt[(size_t)i * nbins + b]; c > 0; --c) h = (float)((double)h + npsqr);
    for (int c = ninv[i]; c > 0; --c) h = (float)((double)h + spread);
  }
  spfh[t] = h;
}

// FPFH step (pfh.cpp:303-333) for long rows: one warp per query, lanes across the bins
__global__ void __launch_bounds__(256) pfh_combine_average_kernel(const float* __restrict__ spfh, const int* __restrict__ list_off,
                                                                  const int2* __restrict__ list, const int* __restrict__ kcount,
                                                                  int n_valid, int nbins, int stride, float* __restrict__ out) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= n_valid) return;
  const int first = list_off[q], count = max(kcount[q] - 1, 0);  // the query itself is not listed (:317)
  double sum_weight = 0.0;
  for (int e = 0; e < count; ++e) sum_weight += 1.0 / (double)__int_as_float(list[first + e].y);
  for (int b0 = 0; b0 < stride; b0 += 8 * kWarp) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int e = 0; e < count; ++e) {
      const int2 en = list[first + e];
      const double weight = 1.0 / (double)__int_as_float(en.y);
      const float* hn = spfh + (size_t)en.x * stride;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int b = b0 + j * kWarp + lane;
        if (b < nbins) acc[j] = (float)((double)acc[j] + (double)hn[b] * weight);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j * kWarp + lane;
      if (b < stride) out[(size_t)q * stride + b] = b < nbins ? (float)((double)acc[j] / sum_weight) : 0.f;  // 0 / 0 = NaN without neighbours (:330)
    }
  }
}

// the way back to input order, point-major rows of nbins floats (no differences in the combined mode, :345)
__global__ void __launch_bounds__(256) pfh_combine_finish_kernel(const float* __restrict__ rows, int stride, int nbins,
                                                                 const int* __restrict__ perm, int n_valid, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_valid * nbins) return;
  const int i = (int)(t / nbins), b = (int)(t % nbins);
  const int dst = perm[i];
  if (dst >= 0) out[(size_t)dst * nbins + b] = rows[(size_t)i * stride + b];
}

}  // namespace

static int run_pfh_combined(cab_ctx* ctx, double radius, int max_nn, int quantum, int flags, float* out_host) {
  const float rf = (float)radius;
  const int nfeat = (flags & CAB_PFH_USE_DIST) ? 4 : 3;
  long long nb = 1;
  for (int f = 0; f < nfeat; ++f) nb *= quantum;  // (int) ceil (pow (quantum_, nr_features_)), pfh.cpp:51
  if (nb > 4096) return fail(ctx, CAB_ERR_ARG, "cab_pfh: quantum ^ features = %lld bins, at most 4096", nb);
  const int nbins = (int)nb, stride = (nbins + 3) & ~3;
  const int n = (int)ctx->n, nv = ctx->n_valid;
  cudaStream_t st = ctx->stream;
  const size_t rows = (size_t)std::max(n, 1);
  if (rows * (size_t)stride > ((size_t)1 << 31)) return fail(ctx, CAB_ERR_OOM, "cab_pfh: %d points x %d bins is more than the combined mode handles", n, nbins);
  if (int rc = reserve(ctx, ctx->b_pfh[0], rows * stride * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_pfh[1], rows * stride * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_pfh[2], rows * nbins * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_out4, rows * nbins * sizeof(int))) return rc;          // pair counts
  if (int rc = reserve(ctx, ctx->b_out1a, rows * sizeof(int))) return rc;                  // invalid pairs
  if (int rc = reserve(ctx, ctx->b_out1b, (rows + 1) * sizeof(int))) return rc;            // list offsets
  if (int rc = reserve(ctx, ctx->b_kcount, rows * sizeof(int))) return rc;
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
  a.max_dist = 2 * radius;
  a.spfh = (float*)ctx->b_pfh[0].p;
  a.out = (float*)ctx->b_pfh[1].p;
  a.kcount = (int*)ctx->b_kcount.p;
  PfhCombineArgs ca{};
  ca.cnt = (int*)ctx->b_out4.p;
  ca.ninv = (int*)ctx->b_out1a.p;
  ca.list_off = (const int*)ctx->b_out1b.p;
  // The order of the features in the histogram binning (:113-121): a_, b_, c_, d_ = 3, 0, 2, 1 with the distance, 2, 0, 1 without
  const int slot4[4] = {3, 0, 2, 1}, slot3[4] = {2, 0, 1, 3};
  for (int f = 0; f < 4; ++f) ca.slot[f] = nfeat == 4 ? slot4[f] : slot3[f];
  const int np = a.p1 - a.p0;
  const float* result = a.spfh;
  CAB_CUDA(ctx, cudaMemsetAsync(ca.cnt, 0, rows * nbins * sizeof(int), st));
  CAB_CUDA(ctx, cudaMemsetAsync(ca.ninv, 0, rows * sizeof(int), st));
  CAB_CUDA(ctx, cudaMemsetAsync(a.kcount, 0, rows * sizeof(int), st));
  if (np > 0 && nv > 0) {
    const unsigned blocks = (np + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (use_thr) pfh_combine_kernel<true, false><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a, ca);
    else pfh_combine_kernel<false, false><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a, ca);
    CAB_LAUNCH_CHECK(ctx);
    const long long elems = (long long)nv * stride;
    pfh_combine_fill_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(ca.cnt, ca.ninv, a.kcount, nv, nbins, stride, a.spfh);
    CAB_LAUNCH_CHECK(ctx);
    if (flags & CAB_PFH_AVERAGE) {
      size_t tmp = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int*)nullptr, (int*)nullptr, nv + 1, st);
      if (int rc = reserve(ctx, ctx->b_cubtmp, tmp + 16)) return rc;
      CAB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_cubtmp.p, tmp, (const int*)a.kcount, (int*)ctx->b_out1b.p, nv + 1, st));
      int total = 0;
      CAB_CUDA(ctx, cudaMemcpyAsync(&total, (const int*)ctx->b_out1b.p + nv, 4, cudaMemcpyDeviceToHost, st));
      CAB_CUDA(ctx, cudaStreamSynchronize(st));
      if (int rc = reserve(ctx, ctx->b_keys[2], (size_t)std::max(total, 1) * sizeof(int2))) return rc;
      ca.list = (int2*)ctx->b_keys[2].p;
      if (use_thr) pfh_combine_kernel<true, true><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a, ca);
      else pfh_combine_kernel<false, true><<<blocks, kWarpsPerBlock * kWarp, 0, st>>>(a, ca);
      CAB_LAUNCH_CHECK(ctx);
      pfh_combine_average_kernel<<<(unsigned)(((long long)nv * kWarp + 255) / 256), 256, 0, st>>>(a.spfh, ca.list_off, ca.list, a.kcount, nv,
                                                                                                  nbins, stride, a.out);
      CAB_LAUNCH_CHECK(ctx);
      result = a.out;
    }
  }
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_pfh[2].p, 0, rows * nbins * sizeof(float), st));
  if (nv > 0) {
    const long long elems = (long long)nv * nbins;
    pfh_combine_finish_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(result, stride, nbins, (const int*)ctx->b_perm.p, nv,
                                                                              (float*)ctx->b_pfh[2].p);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  if (out_host && n > 0)
    CAB_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->b_pfh[2].p, (size_t)n * nbins * sizeof(float), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.pfh_ms, ctx->ev[2], ctx->ev[3]));
  return CAB_OK;
}

int run_pfh(cab_ctx* ctx, double radius, int max_nn, int quantum, int flags, float* out_host) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_pfh: build the grid first");
  if (!ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "cab_pfh: missing normals");
  if (ctx->shard_world > 1) return fail(ctx, CAB_ERR_STATE, "cab_pfh: the averaging pass reads every neighbour's histogram; not available on a shard");
  const float rf = (float)radius;
  if (!(rf > 0.f) || rf > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_pfh: radius %g exceeds the grid cell %g", radius, (double)ctx->cell);
  const int nfeat = (flags & CAB_PFH_USE_DIST) ? 4 : 3;
  const bool combine = (flags & CAB_PFH_COMBINE) != 0;
  if (quantum < 1) return fail(ctx, CAB_ERR_ARG, "cab_pfh: quantum must be >= 1");
  if (combine) return run_pfh_combined(ctx, radius, max_nn, quantum, flags, out_host);
  const int nbins = quantum * nfeat;
  if (nbins > kPfhMaxBins) return fail(ctx, CAB_ERR_ARG, "cab_pfh: quantum * features must be in [1, %d]", kPfhMaxBins);
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
