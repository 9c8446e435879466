// cab_sac.cu -- MSAC plane fit: the table-plane step in front of the object clustering.
//
// Replaces fitSACPlane (cloud_tools/src/table_object_detector_passive.cpp:621-659, same code in
// table_object_detector_sr.cpp): sample_consensus::MSAC over sample_consensus::SACModelPlane with setMaxIterations(500),
// setProbability(0.99), then computeCoefficients, refineCoefficients (least squares over the inliers),
// selectWithinDistance and projectPointsInPlace.  Both classes live in point_cloud_mapping [EXTERNAL, not in the tree]; the
// algorithm restated here is the published one (Torr & Zisserman's M-estimator SAC as that library and later PCL's
// msac.hpp implement it): hypotheses from three sampled points, penalty = sum over the points of min(distance,
// threshold), the best penalty wins, the iteration bound k = log(1 - p) / log(1 - w^3) follows the best inlier ratio w.
//
// The hypothesis loop is sequential upstream (k depends on the best model so far); the penalty of one hypothesis is a
// sum over all points.  Here a batch of 32 hypotheses is scored by one launch (grid = point chunks x hypotheses, fp64
// distances, block partials summed in a fixed order), and the host replays the sequential update rule over the batch in
// order: the chosen model, the iteration count and the inliers are those of the sequential loop on the same sample
// sequence.  The sample sequence is the caller's (the reference's rand() stream cannot be reproduced): `triples`.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include <cub/device/device_select.cuh>

#include "cab_internal.cuh"

namespace cab {
namespace {

constexpr int kSacBatch = 32;
constexpr int kSacChunk = 4096;  // points per block (256 threads x 16)

struct Plane {
  double a, b, c, d;
};

__device__ __forceinline__ double plane_dist(const Plane& m, float x, float y, float z) {
  // fabs(a x + b y + c z + d), double, left to right, no contraction (SACModelPlane::getDistancesToModel)
  return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m.a, (double)x), __dmul_rn(m.b, (double)y)), __dmul_rn(m.c, (double)z)), m.d));
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;  // valid in thread 0
}

// penalty and inlier count of hypothesis blockIdx.y over the points of chunk blockIdx.x
__global__ void __launch_bounds__(256) msac_penalty_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ idx,
                                                           int n, const Plane* __restrict__ models, double thr,
                                                           double* __restrict__ part_pen, int* __restrict__ part_cnt) {
  __shared__ double sh[8];
  const Plane m = models[blockIdx.y];
  const int base = blockIdx.x * kSacChunk;
  double pen = 0;
  int cnt = 0;
  for (int i = base + threadIdx.x; i < min(base + kSacChunk, n); i += blockDim.x) {
    const float* p = xyz + (size_t)(idx ? idx[i] : i) * stride;
    const double d = plane_dist(m, p[0], p[1], p[2]);
    pen += fmin(d, thr);
    cnt += d <= thr ? 1 : 0;
  }
  const double tp = block_sum(pen, sh);
  const double tc = block_sum((double)cnt, sh);
  if (threadIdx.x == 0) {
    part_pen[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tp;
    part_cnt[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = (int)tc;
  }
}

// inlier flags under one model; pass 1 of the refinement: sums of the inliers' coordinates
__global__ void __launch_bounds__(256) sac_flag_sum_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ idx, int n,
                                                           Plane m, double thr, unsigned char* __restrict__ flag,
                                                           double* __restrict__ part /* [blocks][4]: sx sy sz count */) {
  __shared__ double sh[8];
  const int base = blockIdx.x * kSacChunk;
  double sx = 0, sy = 0, sz = 0, c = 0;
  for (int i = base + threadIdx.x; i < min(base + kSacChunk, n); i += blockDim.x) {
    const float* p = xyz + (size_t)(idx ? idx[i] : i) * stride;
    const bool in = plane_dist(m, p[0], p[1], p[2]) <= thr;
    flag[i] = in ? 1 : 0;
    if (in) {
      sx += (double)p[0];
      sy += (double)p[1];
      sz += (double)p[2];
      c += 1.0;
    }
  }
  const double a = block_sum(sx, sh), b = block_sum(sy, sh), d = block_sum(sz, sh), e = block_sum(c, sh);
  if (threadIdx.x == 0) {
    part[4 * (size_t)blockIdx.x] = a;
    part[4 * (size_t)blockIdx.x + 1] = b;
    part[4 * (size_t)blockIdx.x + 2] = d;
    part[4 * (size_t)blockIdx.x + 3] = e;
  }
}

// pass 2: the inliers' scatter about their centroid (cloud_geometry::nearest::computePointNormal's covariance)
__global__ void __launch_bounds__(256) sac_cov_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ idx, int n,
                                                      const unsigned char* __restrict__ flag, double cx, double cy, double cz,
                                                      double* __restrict__ part /* [blocks][6] */) {
  __shared__ double sh[8];
  const int base = blockIdx.x * kSacChunk;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int i = base + threadIdx.x; i < min(base + kSacChunk, n); i += blockDim.x) {
    if (!flag[i]) continue;
    const float* p = xyz + (size_t)(idx ? idx[i] : i) * stride;
    const double dx = (double)p[0] - cx, dy = (double)p[1] - cy, dz = (double)p[2] - cz;
    s[0] += dx * dx;
    s[1] += dx * dy;
    s[2] += dx * dz;
    s[3] += dy * dy;
    s[4] += dy * dz;
    s[5] += dz * dz;
  }
  for (int k = 0; k < 6; ++k) {
    const double t = block_sum(s[k], sh);
    if (threadIdx.x == 0) part[6 * (size_t)blockIdx.x + k] = t;
  }
}

// the selected points (positions into the caller's index list), in order, and their projections onto the plane
__global__ void __launch_bounds__(256) sac_project_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ idx,
                                                          const int* __restrict__ sel, int m, Plane pl, int* __restrict__ out_idx,
                                                          float* __restrict__ out_xyz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int pos = sel[i];
  const int j = idx ? idx[pos] : pos;
  out_idx[i] = j;
  const float* p = xyz + (size_t)j * stride;
  // SACModelPlane::projectPointsInPlace: p -= (a x + b y + c z + d) * n, in double, stored as float
  const double dist = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(pl.a, (double)p[0]), __dmul_rn(pl.b, (double)p[1])), __dmul_rn(pl.c, (double)p[2])), pl.d);
  out_xyz[3 * (size_t)i] = (float)((double)p[0] - dist * pl.a);
  out_xyz[3 * (size_t)i + 1] = (float)((double)p[1] - dist * pl.b);
  out_xyz[3 * (size_t)i + 2] = (float)((double)p[2] - dist * pl.c);
}

__global__ void iota_kernel(int* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// 3x3 symmetric eigen-decomposition (cyclic Jacobi, double): eigenvector of the smallest eigenvalue
void smallest_eigenvector(const double a[6], double n[3]) {
  double m[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    if (std::fabs(m[0][1]) + std::fabs(m[0][2]) + std::fabs(m[1][2]) == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (m[p][q] == 0.0) continue;
        const double theta = (m[q][q] - m[p][p]) / (2.0 * m[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c, apq = m[p][q];
        m[p][p] -= t * apq;
        m[q][q] += t * apq;
        m[p][q] = m[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = m[r][p], arq = m[r][q];
        m[r][p] = m[p][r] = c * arp - s * arq;
        m[r][q] = m[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int c = 1; c < 3; ++c)
    if (m[c][c] < m[best][best]) best = c;
  const double len = std::sqrt(v[0][best] * v[0][best] + v[1][best] * v[1][best] + v[2][best] * v[2][best]);
  for (int r = 0; r < 3; ++r) n[r] = v[r][best] / len;
}

// SACModelPlane::computeModelCoefficients: the plane through three points; false when they are collinear
bool plane_from_triple(const float* p0, const float* p1, const float* p2, Plane* out) {
  const double ux = (double)p1[0] - p0[0], uy = (double)p1[1] - p0[1], uz = (double)p1[2] - p0[2];
  const double vx = (double)p2[0] - p0[0], vy = (double)p2[1] - p0[1], vz = (double)p2[2] - p0[2];
  double a = uy * vz - uz * vy, b = uz * vx - ux * vz, c = ux * vy - uy * vx;
  const double len = std::sqrt(a * a + b * b + c * c);
  if (!(len > 0.0) || !std::isfinite(len)) return false;
  a /= len;
  b /= len;
  c /= len;
  out->a = a;
  out->b = b;
  out->c = c;
  out->d = -(a * (double)p0[0] + b * (double)p0[1] + c * (double)p0[2]);
  return true;
}

}  // namespace
}  // namespace cab

using namespace cab;

extern "C" int64_t cab_fit_plane_msac(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride, const int32_t* indices, int64_t n_idx,
                                      double threshold, int32_t max_iterations, double probability, const int32_t* triples,
                                      int64_t n_triples, double coeff[4], int32_t* inliers, float* projected_xyz, int64_t cap,
                                      int32_t* iterations_run, int32_t* best_iteration) {
  if (!ctx) return CAB_ERR_ARG;
  if (!xyz || n < 0 || stride < 3 || !coeff) return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: bad cloud / coeff");
  if (indices && n_idx < 0) return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: bad index list");
  if (!(threshold > 0) || max_iterations < 1 || !(probability > 0 && probability < 1))
    return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: threshold > 0, max_iterations >= 1, 0 < probability < 1");
  if (!triples || n_triples < 1) return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: the sample sequence (triples) is the caller's");
  const int m = (int)(indices ? n_idx : n);
  if (iterations_run) *iterations_run = 0;
  if (best_iteration) *best_iteration = -1;
  for (int k = 0; k < 4; ++k) coeff[k] = 0.0;
  if (m < 3) return 0;
  for (int64_t t = 0; t < 3 * n_triples; ++t)
    if (triples[t] < 0 || triples[t] >= m) return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: sample position %d outside the %d points", triples[t], m);
  if (indices)
    for (int64_t i = 0; i < n_idx; ++i)
      if (indices[i] < 0 || indices[i] >= n) return fail(ctx, CAB_ERR_ARG, "cab_fit_plane_msac: index %d outside the cloud", indices[i]);
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // ---- the points (and the index list) on the device ----------------------------------------------------
  const size_t cloud_bytes = (size_t)n * stride * sizeof(float);
  if (int rc = reserve(ctx, ctx->b_pfh[0], cloud_bytes + 64)) return rc;
  if (int rc = reserve(ctx, ctx->b_pfh[1], (size_t)m * 4 * 3 + 64)) return rc;  // index list, selection, selected
  const int blocks = (m + kSacChunk - 1) / kSacChunk;
  const size_t part_doubles = (size_t)kSacBatch * blocks + 6 * (size_t)blocks + 64;
  if (int rc = reserve(ctx, ctx->b_pfh[2], part_doubles * 8 + (size_t)kSacBatch * blocks * 4 + kSacBatch * sizeof(Plane) + (size_t)m + 256)) return rc;
  float* d_xyz = (float*)ctx->b_pfh[0].p;
  int* d_idx = indices ? (int*)ctx->b_pfh[1].p : nullptr;
  int* d_iota = (int*)ctx->b_pfh[1].p + m;
  int* d_sel = d_iota + m;
  double* d_part = (double*)ctx->b_pfh[2].p;
  int* d_cnt = (int*)(d_part + part_doubles);
  Plane* d_models = (Plane*)(d_cnt + (size_t)kSacBatch * blocks);
  unsigned char* d_flag = (unsigned char*)(d_models + kSacBatch);
  CAB_CUDA(ctx, cudaMemcpyAsync(d_xyz, xyz, cloud_bytes, cudaMemcpyHostToDevice, st));
  if (indices) CAB_CUDA(ctx, cudaMemcpyAsync(d_idx, indices, (size_t)m * 4, cudaMemcpyHostToDevice, st));
  auto point = [&](int pos) { return xyz + (size_t)(indices ? indices[pos] : pos) * stride; };

  // ---- MSAC: batches of hypotheses scored on the device, the sequential update rule replayed on the host -------
  std::vector<Plane> models(kSacBatch);
  std::vector<char> valid(kSacBatch);
  std::vector<double> h_pen((size_t)kSacBatch * blocks);
  std::vector<int> h_cnt((size_t)kSacBatch * blocks);
  double best_penalty = DBL_MAX, k = 1.0;
  Plane best{};
  int best_it = -1, iterations = 0;
  int64_t smp = 0;  // next entry of the sample sequence
  bool done = false;
  while (!done && iterations < k && smp < n_triples) {
    const int nb = (int)std::min<int64_t>(kSacBatch, n_triples - smp);
    for (int h = 0; h < nb; ++h) {
      const int32_t* t = triples + 3 * (size_t)(smp + h);
      valid[h] = (t[0] != t[1] && t[0] != t[2] && t[1] != t[2]) && plane_from_triple(point(t[0]), point(t[1]), point(t[2]), &models[h]);
      if (!valid[h]) models[h] = Plane{0, 0, 0, 0};
    }
    CAB_CUDA(ctx, cudaMemcpyAsync(d_models, models.data(), nb * sizeof(Plane), cudaMemcpyHostToDevice, st));
    msac_penalty_kernel<<<dim3(blocks, nb), 256, 0, st>>>(d_xyz, stride, d_idx, m, d_models, threshold, d_part, d_cnt);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cudaMemcpyAsync(h_pen.data(), d_part, (size_t)nb * blocks * 8, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)nb * blocks * 4, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    for (int h = 0; h < nb; ++h, ++smp) {
      if (!(iterations < k)) {  // the sequential loop would have stopped before this hypothesis
        done = true;
        break;
      }
      if (!valid[h]) continue;  // getSamples draws again: a degenerate triple is not an iteration
      {
        double pen = 0;
        long long cnt = 0;
        for (int b = 0; b < blocks; ++b) {
          pen += h_pen[(size_t)h * blocks + b];
          cnt += h_cnt[(size_t)h * blocks + b];
        }
        if (pen < best_penalty) {
          best_penalty = pen;
          best = models[h];
          best_it = (int)(smp);
          const double w = (double)cnt / (double)m;
          double p_no_outliers = 1.0 - std::pow(w, 3.0);
          p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
          p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
          k = std::log(1.0 - probability) / std::log(p_no_outliers);
        }
      }
      iterations += 1;
      if (iterations > max_iterations) {
        done = true;
        break;
      }
    }
  }
  if (iterations_run) *iterations_run = iterations;
  if (best_iteration) *best_iteration = best_it;
  if (best_it < 0) return 0;  // no valid hypothesis (computeModel returns false)

  // ---- refineCoefficients: least-squares plane of the best model's inliers ---------------------------------
  std::vector<double> h_part(6 * (size_t)blocks);
  sac_flag_sum_kernel<<<blocks, 256, 0, st>>>(d_xyz, stride, d_idx, m, best, threshold, d_flag, d_part);
  CAB_LAUNCH_CHECK(ctx);
  CAB_CUDA(ctx, cudaMemcpyAsync(h_part.data(), d_part, 4 * (size_t)blocks * 8, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  double sx = 0, sy = 0, sz = 0, cnt = 0;
  for (int b = 0; b < blocks; ++b) {
    sx += h_part[4 * (size_t)b];
    sy += h_part[4 * (size_t)b + 1];
    sz += h_part[4 * (size_t)b + 2];
    cnt += h_part[4 * (size_t)b + 3];
  }
  Plane refined = best;
  if (cnt >= 3) {
    const double cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
    sac_cov_kernel<<<blocks, 256, 0, st>>>(d_xyz, stride, d_idx, m, d_flag, cx, cy, cz, d_part);
    CAB_LAUNCH_CHECK(ctx);
    CAB_CUDA(ctx, cudaMemcpyAsync(h_part.data(), d_part, 6 * (size_t)blocks * 8, cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaStreamSynchronize(st));
    double cov[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < blocks; ++b)
      for (int q = 0; q < 6; ++q) cov[q] += h_part[6 * (size_t)b + q];
    double nrm[3];
    smallest_eigenvector(cov, nrm);
    if (std::isfinite(nrm[0]) && std::isfinite(nrm[1]) && std::isfinite(nrm[2])) {
      // the eigenvector's sign is arbitrary upstream; here it follows the sampled model's side
      if (nrm[0] * best.a + nrm[1] * best.b + nrm[2] * best.c < 0)
        for (double& v : nrm) v = -v;
      refined = Plane{nrm[0], nrm[1], nrm[2], -(nrm[0] * cx + nrm[1] * cy + nrm[2] * cz)};
    }
  }
  coeff[0] = refined.a;
  coeff[1] = refined.b;
  coeff[2] = refined.c;
  coeff[3] = refined.d;

  // ---- selectWithinDistance under the refined model, in list order; projectPointsInPlace ----------------------
  sac_flag_sum_kernel<<<blocks, 256, 0, st>>>(d_xyz, stride, d_idx, m, refined, threshold, d_flag, d_part);
  CAB_LAUNCH_CHECK(ctx);
  iota_kernel<<<(m + 255) / 256, 256, 0, st>>>(d_iota, m);
  CAB_LAUNCH_CHECK(ctx);
  size_t tmp = 0;
  int* d_num = (int*)d_cnt;
  cub::DeviceSelect::Flagged(nullptr, tmp, d_iota, d_flag, d_sel, d_num, m, st);
  if (int rc = reserve(ctx, ctx->b_cubtmp, tmp + 16)) return rc;
  CAB_CUDA(ctx, cub::DeviceSelect::Flagged(ctx->b_cubtmp.p, tmp, d_iota, d_flag, d_sel, d_num, m, st));
  ctx->tm.kernel_launches += 2;
  int n_in = 0;
  CAB_CUDA(ctx, cudaMemcpyAsync(&n_in, d_num, 4, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  if (n_in == 0 || (!inliers && !projected_xyz) || n_in > cap) return n_in;
  if (int rc = reserve(ctx, ctx->b_out4, (size_t)n_in * 16 + 64)) return rc;
  int* d_out_idx = (int*)ctx->b_out4.p;
  float* d_out_xyz = (float*)(d_out_idx + n_in);
  sac_project_kernel<<<(n_in + 255) / 256, 256, 0, st>>>(d_xyz, stride, d_idx, d_sel, n_in, refined, d_out_idx, d_out_xyz);
  CAB_LAUNCH_CHECK(ctx);
  if (inliers) CAB_CUDA(ctx, cudaMemcpyAsync(inliers, d_out_idx, (size_t)n_in * 4, cudaMemcpyDeviceToHost, st));
  if (projected_xyz) CAB_CUDA(ctx, cudaMemcpyAsync(projected_xyz, d_out_xyz, (size_t)n_in * 12, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return n_in;
}
