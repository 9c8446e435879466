// cab_rsd.cu -- Radius-based Surface Descriptor for every point of the cloud.
// Replaces HOT LOOP 1 + HOT LOOP 2 of cloud_algos::LocalRadiusEstimation::process
// (cloud_algos/src/radius_estimation.cpp:118-124 and :140-202): no neighbour list is ever
// materialised; each query keeps per-distance-bin extreme cosines in shared memory while the
// candidate tiles stream past, then solves the two one-parameter least-squares fits.
//
// Arithmetic notes (DESIGN.md "RSD kernel"):
//  * cosine is the reference's fp32 expression (nx*nx' + ny*ny') + nz*nz' (:153-155), not contracted.
//  * angle = acos(cosine) folded at pi/2 (:160-161) is monotone in |cosine|, so per bin only the
//    cosines of extreme |value| are tracked; acos is evaluated 2*ndiv times per query, not per pair.
//  * the distance bin floor(ndiv*sqrt((double)d2)/radius) (:165-168) is evaluated through exact
//    fp32 d2 thresholds computed on the host with the reference's double expression, clamped to
//    ndiv-1 (the reference indexes out of bounds at dist == radius).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

constexpr int kMaxDiv = 64;

struct RsdArgs {
  GridView g;
  int p0, p1;
  float r, r2;
  const float4* nrm;       // sorted order
  float2* out;             // sorted order (r_min, r_max)
  const float* thr_d2;     // optional max_nn thresholds
  const int* thr_idx;
  const float* bin_thr;    // ndiv + 1 fp32 d2 thresholds, bin_thr[0] = -inf, bin_thr[ndiv] = +inf
  int ndiv;
  int flags;
  float bin_scale;         // ndiv / radius
  double radius, plane_radius;
  unsigned long long* stats;
};

// angle between the two lines, radius_estimation.cpp:158-161
template <bool kExact>
__device__ __forceinline__ double fold_angle(float c) {
  if (kExact) {
    double a = acos((double)c);
    if (a > M_PI / 2) a = M_PI - a;
    return a;
  } else {
    return (double)acosf(fabsf(c));
  }
}

template <bool kExact>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) rsd_kernel(const RsdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* ptile = reinterpret_cast<float4*>(smem_raw);                      // [W][32]
  float4* ntile = ptile + kWarpsPerBlock * kWarp;                           // [W][32]
  float2* bins = reinterpret_cast<float2*>(ntile + kWarpsPerBlock * kWarp);  // [W][ndiv][32]
  float* thr = reinterpret_cast<float*>(bins + kWarpsPerBlock * a.ndiv * kWarp);  // [ndiv+1]
  __shared__ unsigned long long blk_stats[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ndiv = a.ndiv;
  for (int i = threadIdx.x; i <= ndiv; i += blockDim.x) thr[i] = a.bin_thr[i];
  if (threadIdx.x < 2) blk_stats[threadIdx.x] = 0;
  __syncthreads();
  const int pid = a.p0 + blockIdx.x * kWarpsPerBlock + warp;
  if (pid < a.p1) {
    const GridView& g = a.g;
    const PacketCtx pc = load_packet(g, pid, lane, a.r);
    const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
    const float4 nq = a.nrm[pc.qi];
    const float r2 = a.r2;
    const bool use_thr = a.thr_d2 != nullptr;
    float td2 = 0.f;
    int tidx = 0;
    if (use_thr) {
      td2 = a.thr_d2[pc.qi];
      tidx = a.thr_idx[pc.qi];
    }
    float4* my_p = ptile + warp * kWarp;
    float4* my_n = ntile + warp * kWarp;
    float2* my_b = bins + (size_t)warp * ndiv * kWarp + lane;  // bin b at my_b[b * 32]
    // .x: cosine of smallest |value| (largest angle), .y: cosine of largest |value| (smallest angle)
    for (int b = 0; b < ndiv; ++b) my_b[b * kWarp] = make_float2(INFINITY, 0.f);
    if (a.flags & CAB_RSD_SEED_BIN0) my_b[0] = make_float2(1.f, 1.f);
    int k = 0;
    const float bscale = a.bin_scale;
    const int tested = for_each_chunk(
        g, pc, lane,
        [&](int j, bool valid, const float4& c) {
          my_p[lane] = c;
          my_n[lane] = valid ? a.nrm[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        },
        [&](int base, int cnt) {
          const int cnt4 = (cnt + 3) & ~3;
#pragma unroll 4
          for (int m = 0; m < cnt4; ++m) {
            const float4 c = my_p[m];
            const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            bool hit = d2 <= r2;
            if (use_thr && hit) hit = d2 < td2 || (d2 == td2 && g.perm[base + m] <= tidx);
            if (hit) {
              ++k;
              if (base + m != pc.qi) {  // the query itself is skipped (:150 starts at ni = 1)
                const float4 nm = my_n[m];
                float cs = __fadd_rn(__fadd_rn(__fmul_rn(nq.x, nm.x), __fmul_rn(nq.y, nm.y)), __fmul_rn(nq.z, nm.z));
                if (cs > 1.f) cs = 1.f;  // :158-159, NaN falls through
                if (cs < -1.f) cs = -1.f;
                const float sq = d2 > 0.f ? d2 * rsqrtf(d2) : 0.f;
                int b = min((int)(sq * bscale), ndiv - 1);
                if (d2 < thr[b]) --b;
                else if (d2 >= thr[b + 1]) ++b;
                float2 v = my_b[b * kWarp];
                const float ac = fabsf(cs);
                if (ac < fabsf(v.x)) v.x = cs;
                if (ac >= fabsf(v.y)) v.y = cs;
                my_b[b * kWarp] = v;
              }
            }
          }
        });

    // ---- least-squares fit of the min / max angle lines, radius_estimation.cpp:175-202 ----
    double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
    for (int di = 0; di < ndiv; ++di) {
      const float2 v = my_b[di * kWarp];
      if (fabsf(v.x) <= 1.f) {  // bin not empty (:181)
        const double p_min = fold_angle<kExact>(v.y), p_max = fold_angle<kExact>(v.x);
        const double f = (di + 0.5) * a.radius / ndiv;
        Amint_Amin = __dadd_rn(Amint_Amin, __dmul_rn(p_min, p_min));
        Amint_d = __dadd_rn(Amint_d, __dmul_rn(p_min, f));
        Amaxt_Amax = __dadd_rn(Amaxt_Amax, __dmul_rn(p_max, p_max));
        Amaxt_d = __dadd_rn(Amaxt_d, __dmul_rn(p_max, f));
      }
    }
    double max_radius = (Amint_Amin == 0) ? a.plane_radius : fmin(Amint_d / Amint_Amin, a.plane_radius);
    double min_radius = (Amaxt_Amax == 0) ? a.plane_radius : fmin(Amaxt_d / Amaxt_Amax, a.plane_radius);
    float rmin = (float)min_radius, rmax = (float)max_radius;
    if (a.flags & CAB_RSD_SCALE_SORT) {
      const float x = rmax * 1.1f, y = rmin * 0.9f;
      rmin = fminf(x, y);
      rmax = fmaxf(x, y);
    }
    if (pc.active) a.out[pc.qi] = make_float2(rmin, rmax);
    unsigned long long ks = pc.active ? (unsigned long long)k : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) ks += __shfl_xor_sync(kFull, ks, o);
    if (lane == 0) {
      atomicAdd(&blk_stats[0], ks);
      atomicAdd(&blk_stats[1], (unsigned long long)tested * (unsigned)pc.count);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 && blk_stats[threadIdx.x]) atomicAdd(a.stats + threadIdx.x, blk_stats[threadIdx.x]);
}

__global__ void fill_invalid_rsd(float2* out, int begin, int end, float v) {
  int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < end) out[i] = make_float2(v, v);
}

// Smallest fp32 d2 whose reference bin floor(ndiv*sqrt((double)d2)/radius) is >= b.
float bin_threshold(int b, int ndiv, double radius, float r2) {
  auto bin_of = [&](float d2) { return (int)std::floor(ndiv * std::sqrt((double)d2) / radius); };
  if (bin_of(r2) < b) return INFINITY;
  uint32_t lo = 0, hi;  // invariant: bin(lo) < b <= bin(hi); d2 >= 0 so bit patterns are ordered
  std::memcpy(&hi, &r2, 4);
  if (bin_of(0.f) >= b) return 0.f;
  while (hi - lo > 1) {
    uint32_t mid = lo + (hi - lo) / 2;
    float f;
    std::memcpy(&f, &mid, 4);
    if (bin_of(f) >= b) hi = mid; else lo = mid;
  }
  float f;
  std::memcpy(&f, &hi, 4);
  return f;
}

}  // namespace

int run_rsd(cab_ctx* ctx, double r, int max_nn, int ndiv, double plane_radius, int flags) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_rsd: build the grid first");
  if (!ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "cab_rsd: missing normals");
  const float rf = (float)r;
  if (!(rf > 0.f) || rf > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_rsd: radius %g exceeds the grid cell %g", r, (double)ctx->cell);
  if (ndiv < 1 || ndiv > kMaxDiv) return fail(ctx, CAB_ERR_ARG, "cab_rsd: distance_div must be in [1, %d]", kMaxDiv);
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (int rc = reserve(ctx, ctx->b_rsd, (size_t)std::max(n, 1) * sizeof(float2))) return rc;
  if (int rc = reserve(ctx, ctx->b_stats, 64)) return rc;
  const bool use_thr = max_nn > 0;
  if (use_thr)
    if (int rc = run_thresholds(ctx, rf, max_nn)) return rc;
  // bin thresholds
  float thr[kMaxDiv + 1];
  const float r2 = rf * rf;
  thr[0] = -INFINITY;
  for (int b = 1; b < ndiv; ++b) thr[b] = bin_threshold(b, ndiv, r, r2);
  thr[ndiv] = INFINITY;
  if (int rc = reserve(ctx, ctx->b_misc, sizeof(thr))) return rc;
  if (int rc = reserve_pinned(ctx, sizeof(thr))) return rc;
  std::memcpy(ctx->h_pin, thr, sizeof(float) * (ndiv + 1));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_misc.p, ctx->h_pin, sizeof(float) * (ndiv + 1), cudaMemcpyHostToDevice, st));
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_stats.p, 0, 64, st));
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
  RsdArgs a{};
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  a.r = rf;
  a.r2 = r2;
  a.nrm = (const float4*)ctx->b_nrm.p;
  a.out = (float2*)ctx->b_rsd.p;
  a.thr_d2 = use_thr ? (const float*)ctx->b_thr_d2.p : nullptr;
  a.thr_idx = use_thr ? (const int*)ctx->b_thr_idx.p : nullptr;
  a.bin_thr = (const float*)ctx->b_misc.p;
  a.ndiv = ndiv;
  a.flags = flags;
  a.bin_scale = (float)(ndiv / r);
  a.radius = r;
  a.plane_radius = plane_radius;
  a.stats = (unsigned long long*)ctx->b_stats.p;
  const size_t smem = (size_t)kWarpsPerBlock * kWarp * sizeof(float4) * 2 +
                      (size_t)kWarpsPerBlock * ndiv * kWarp * sizeof(float2) + (ndiv + 1) * sizeof(float);
  const int np = a.p1 - a.p0;
  if (np > 0) {
    const unsigned blocks = (np + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (ctx->cfg.exact) {
      CAB_CUDA(ctx, cudaFuncSetAttribute(rsd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rsd_kernel<true><<<blocks, kWarpsPerBlock * kWarp, smem, st>>>(a);
    } else {
      CAB_CUDA(ctx, cudaFuncSetAttribute(rsd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rsd_kernel<false><<<blocks, kWarpsPerBlock * kWarp, smem, st>>>(a);
    }
    CAB_LAUNCH_CHECK(ctx);
  }
  if (n > ctx->n_valid) {
    fill_invalid_rsd<<<(n - ctx->n_valid + 255) / 256, 256, 0, st>>>((float2*)ctx->b_rsd.p, ctx->n_valid, n,
                                                                    (float)plane_radius);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->b_stats.p, 16, cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.rsd_ms, ctx->ev[4], ctx->ev[5]));
  ctx->tm.neighbour_sum = (int64_t)((const unsigned long long*)ctx->h_pin)[0];
  ctx->tm.candidate_sum = (int64_t)((const unsigned long long*)ctx->h_pin)[1];
  ctx->have_rsd = true;
  return CAB_OK;
}

}  // namespace cab
