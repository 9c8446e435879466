// cab_normals.cu -- radius-neighbourhood PCA normals.
// Replaces pcl::NormalEstimation::compute with setRadiusSearch(r) as called at
// color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:76-81 (and exampleRSD.cpp:51-58,
// computeGRSD.cpp:101-106, hough_segmentation/src/rsd.cpp:65-72): neighbours with d2 <= r2
// (self included), covariance about the centroid, eigenvector of the smallest eigenvalue,
// curvature = l0 / (l0+l1+l2), flipped towards the viewpoint; fewer than 3 neighbours -> NaN.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

// Cyclic Jacobi on a symmetric 3x3 (a = xx,xy,xz,yy,yz,zz).  Returns the smallest eigenvalue and
// its eigenvector.  Everything stays in registers (fully unrolled, no dynamic indexing).
template <typename T, int kSweeps>
__device__ __forceinline__ void smallest_eigen(T a00, T a01, T a02, T a11, T a12, T a22, T& lam, T& nx, T& ny,
                                               T& nz) {
  T v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
#define CAB_ROT(app, aqq, apq, arp, arq, v0p, v0q, v1p, v1q, v2p, v2q)          \
  if (apq != T(0)) {                                                           \
    T theta = (aqq - app) / (T(2) * apq);                                      \
    T t = T(1) / (fabs(theta) + sqrt(theta * theta + T(1)));                   \
    t = theta < T(0) ? -t : t;                                                 \
    T c = T(1) / sqrt(t * t + T(1)), s = t * c;                                \
    app -= t * apq;                                                            \
    aqq += t * apq;                                                            \
    apq = T(0);                                                                \
    T x = arp, y = arq;                                                        \
    arp = c * x - s * y;                                                       \
    arq = s * x + c * y;                                                       \
    x = v0p; y = v0q; v0p = c * x - s * y; v0q = s * x + c * y;                \
    x = v1p; y = v1q; v1p = c * x - s * y; v1q = s * x + c * y;                \
    x = v2p; y = v2q; v2p = c * x - s * y; v2q = s * x + c * y;                \
  }
#pragma unroll 1
  for (int sweep = 0; sweep < kSweeps; ++sweep) {
    if (a01 == T(0) && a02 == T(0) && a12 == T(0)) break;
    CAB_ROT(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21)  // (p,q)=(0,1), r=2
    CAB_ROT(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22)  // (0,2), r=1
    CAB_ROT(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22)  // (1,2), r=0
  }
#undef CAB_ROT
  lam = a00; nx = v00; ny = v10; nz = v20;
  if (a11 < lam) { lam = a11; nx = v01; ny = v11; nz = v21; }
  if (a22 < lam) { lam = a22; nx = v02; ny = v12; nz = v22; }
}

struct NormalsArgs {
  GridView g;
  int p0, p1;
  const int* range;       // optional device-side {p0, p1} (own + halo packets of a slab); null: the values above
  float r, r2;
  float vpx, vpy, vpz;
  float4* nrm;              // sorted order
  float4* nrm_in;           // optional: the same normals at their input indices (cab_normals_rsd, input-order layout)
  int* kcount;              // sorted order
  const float* thr_d2;      // optional max_nn thresholds (sorted order), may be null
  const int* thr_idx;
  const unsigned char* done;  // optional (input order): packets whose queries are all done are skipped (k-NN rounds)
  unsigned long long* stats;  // kStatSlots x {neighbour sum, candidate sum}
  // kHist: the d2 histogram of the truncated fast RSD pass that follows at the same radius (cab_topk.cu nn_hist_kernel
  // does the same in a traversal of its own): per query the bin of neighbour max_nn + 1 and how many of that bin's
  // candidates are kept, per packet whether some query's bin holds more candidates than the RSD pass can list
  int* code;
  unsigned char* flag;      // indexed by packet - p0
  int hist_max_nn;
  float hist_scale;         // kTruncBins / r2
  float one;                // 1.0f (see accum_pred)
};

// fp32 accumulate of one candidate, predicated on d2 <= r2 (explicit predication: the compiler's
// branchy version costs three more issue slots per candidate).  The three first-moment sums are written as
// fma(d, 1, s) -- the same sum, the same rounding -- because a loop of FFMAs runs 2 % faster here than the same loop
// with FADDs among them (scripts/ubench/normals_mix.cu: 93.7 against 95.8 cycles per four candidates); `one` arrives as
// a kernel argument so that ptxas cannot fold the multiplication away.
__device__ __forceinline__ void accum_pred(float d2, float r2, float dx, float dy, float dz, float& s1x, float& s1y,
                                           float& s1z, float& sxx, float& sxy, float& sxz, float& syy, float& syz,
                                           float& szz, int& k, float one) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.le.f32 p, %10, %11;\n\t"
      "@p fma.rn.f32 %0, %12, %15, %0;\n\t"
      "@p fma.rn.f32 %1, %13, %15, %1;\n\t"
      "@p fma.rn.f32 %2, %14, %15, %2;\n\t"
      "@p fma.rn.f32 %3, %12, %12, %3;\n\t"
      "@p fma.rn.f32 %4, %12, %13, %4;\n\t"
      "@p fma.rn.f32 %5, %12, %14, %5;\n\t"
      "@p fma.rn.f32 %6, %13, %13, %6;\n\t"
      "@p fma.rn.f32 %7, %13, %14, %7;\n\t"
      "@p fma.rn.f32 %8, %14, %14, %8;\n\t"
      "@p add.s32 %9, %9, 1;\n\t}"
      : "+f"(s1x), "+f"(s1y), "+f"(s1z), "+f"(sxx), "+f"(sxy), "+f"(sxz), "+f"(syy), "+f"(syz), "+f"(szz), "+r"(k)
      : "f"(d2), "f"(r2), "f"(dx), "f"(dy), "f"(dz), "f"(one));
}

// one candidate of the kHist variant: bin of the 64-bin d2 histogram (the expression of nn_hist_kernel, same bits), misses
// to the spare row (an unconditional reduction: ptxas would turn a predicated one into a branch)
__device__ __forceinline__ void hist_add(float d2, float r2, float scale, unsigned col) {
  const int b = d2 <= r2 ? min(__float2int_rz(d2 * scale), kTruncBins - 1) : kTruncBins;
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(col + 128u * (unsigned)b));
}

template <bool kExact, bool kUseThr, bool kHist = false>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) normals_kernel(const NormalsArgs a) {
  __shared__ ChunkTile tiles[kWarpsPerBlock];
  extern __shared__ unsigned hist_dyn[];  // kHist: [warps][kTruncBins + 1][32], row kTruncBins: the misses (dynamic: 66.5 KB)
  typedef unsigned HistRows[kTruncBins + 1][kWarp];
  HistRows* hist = reinterpret_cast<HistRows*>(hist_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned hcol = kHist ? (unsigned)__cvta_generic_to_shared(&hist[warp][0][lane]) : 0u;
  const float hscale = a.hist_scale;
  const float one = a.one;
  const GridView& g = a.g;
  const int p0 = a.range ? a.range[0] : a.p0, p1 = a.range ? a.range[1] : a.p1;
  for (;;) {
  const int pid = p0 + next_packet(a.stats, lane);
  if (pid >= p1) break;
  ChunkTile* tile = &tiles[warp];
  const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
  if (a.done && !__any_sync(kFull, pc.active && !a.done[g.perm[pc.qi]])) continue;
  const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
  const float r2 = a.r2;
  using Acc = typename std::conditional<kExact, double, float>::type;
  Acc s1x = 0, s1y = 0, s1z = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
  int k = 0;
  int tested;
  if constexpr (!kExact && !kUseThr) {
    // fast path: packed fp32x2 distance test, predicated fp32 accumulation
    const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
    if constexpr (kHist) {
      for (int b = 0; b <= kTruncBins; ++b) hist[warp][b][lane] = 0;
      __syncwarp();
    }
    tested = for_each_chunk(g, pc, lane, tile, [&](int cnt, const float4&, int, bool) {
      const float4* tx = reinterpret_cast<const float4*>(tile->x);
      const float4* ty = reinterpret_cast<const float4*>(tile->y);
      const float4* tz = reinterpret_cast<const float4*>(tile->z);
      const int groups = (cnt + 3) >> 2;
#pragma unroll 2
      for (int g4 = 0; g4 < groups; ++g4) {
        const float4 X = tx[g4], Y = ty[g4], Z = tz[g4];
        f32x2 dx = sub2(pack2(X.x, X.y), qx2), dy = sub2(pack2(Y.x, Y.y), qy2), dz = sub2(pack2(Z.x, Z.y), qz2);
        f32x2 d2 = add2(add2(sq2(dx), sq2(dy)), sq2(dz));
        float d2a, d2b, xa, xb, ya, yb, za, zb;
        unpack2(d2, d2a, d2b); unpack2(dx, xa, xb); unpack2(dy, ya, yb); unpack2(dz, za, zb);
        accum_pred(d2a, r2, xa, ya, za, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, k, one);
        accum_pred(d2b, r2, xb, yb, zb, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, k, one);
        if constexpr (kHist) {
          hist_add(d2a, r2, hscale, hcol);
          hist_add(d2b, r2, hscale, hcol);
        }
        dx = sub2(pack2(X.z, X.w), qx2); dy = sub2(pack2(Y.z, Y.w), qy2); dz = sub2(pack2(Z.z, Z.w), qz2);
        d2 = add2(add2(sq2(dx), sq2(dy)), sq2(dz));
        unpack2(d2, d2a, d2b); unpack2(dx, xa, xb); unpack2(dy, ya, yb); unpack2(dz, za, zb);
        accum_pred(d2a, r2, xa, ya, za, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, k, one);
        accum_pred(d2b, r2, xb, yb, zb, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, k, one);
        if constexpr (kHist) {
          hist_add(d2a, r2, hscale, hcol);
          hist_add(d2b, r2, hscale, hcol);
        }
      }
    });
    if constexpr (kHist) {
      // the bin that holds neighbour number max_nn + 1 (the first one dropped), as nn_hist_kernel finds it
      __syncwarp();
      asm volatile("" ::: "memory");
      int kk = 0, bin = 255, before = 0, in_bin = 0;
      for (int b = 0; b < kTruncBins; ++b) {
        const int c = (int)hist[warp][b][lane];
        if (bin == 255 && kk + c >= a.hist_max_nn + 1) {
          bin = b;
          before = kk;
          in_bin = c;
        }
        kk += c;
      }
      // (the RSD pass lists the bin's candidates by their 16-bit position in the candidate stream)
      const bool over = pc.active && bin != 255 && (in_bin > kTruncCap || pc.total > 65535);
      if (pc.active) a.code[pc.qi] = bin == 255 ? 255 : (bin | ((a.hist_max_nn - before) << 8));
      if (__any_sync(kFull, over) && lane == 0) a.flag[pid - a.p0] = 1;
      __syncwarp();
    }
  } else {
    // exact / truncated path: hit mask per chunk, then only the hits are visited; the candidate is
    // fetched from the lane that staged it (warp shuffle, no shared-memory bank conflicts)
    float td2 = INFINITY;
    int tidx = INT_MAX;
    if (kUseThr) {
      td2 = a.thr_d2[pc.qi];
      tidx = a.thr_idx[pc.qi];
    }
    tested = for_each_chunk(g, pc, lane, tile, [&](int, const float4& c, int, bool) {
      unsigned mask = chunk_hit_mask(tile, qx, qy, qz, r2);
      const int iters = __reduce_max_sync(kFull, __popc(mask));
      for (int it = 0; it < iters; ++it) {
        const bool has = mask != 0;
        const int m = __ffs(mask) - 1;
        mask &= mask - 1;
        const float cx = __shfl_sync(kFull, c.x, m), cy = __shfl_sync(kFull, c.y, m), cz = __shfl_sync(kFull, c.z, m);
        if (has) {
          const float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
          bool hit = true;
          if (kUseThr) {
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            hit = d2 < td2 || (d2 == td2 && g.perm[tile->idx[m]] <= tidx);
          }
          if (hit) {
            if constexpr (kExact) {
              const double ex = dx, ey = dy, ez = dz;
              s1x += ex; s1y += ey; s1z += ez;
              sxx = __fma_rn(ex, ex, sxx); sxy = __fma_rn(ex, ey, sxy); sxz = __fma_rn(ex, ez, sxz);
              syy = __fma_rn(ey, ey, syy); syz = __fma_rn(ey, ez, syz); szz = __fma_rn(ez, ez, szz);
            } else {
              s1x += dx; s1y += dy; s1z += dz;
              sxx = fmaf(dx, dx, sxx); sxy = fmaf(dx, dy, sxy); sxz = fmaf(dx, dz, sxz);
              syy = fmaf(dy, dy, syy); syz = fmaf(dy, dz, syz); szz = fmaf(dz, dz, szz);
            }
            ++k;
          }
        }
      }
    });
  }

  // ---- finalise: covariance about the centroid (fp64), eigen-solve ---------------------
  float4 out;
  if (k < 3) {
    const float nan = __int_as_float(0x7fc00000);
    out = make_float4(nan, nan, nan, nan);
  } else {
    // written with explicit roundings so the CPU oracle's plain expressions give the same bits
    const double inv = __drcp_rn((double)k);
    const double mx = __dmul_rn((double)s1x, inv), my = __dmul_rn((double)s1y, inv), mz = __dmul_rn((double)s1z, inv);
    const double cxx = __dsub_rn(__dmul_rn((double)sxx, inv), __dmul_rn(mx, mx));
    const double cxy = __dsub_rn(__dmul_rn((double)sxy, inv), __dmul_rn(mx, my));
    const double cxz = __dsub_rn(__dmul_rn((double)sxz, inv), __dmul_rn(mx, mz));
    const double cyy = __dsub_rn(__dmul_rn((double)syy, inv), __dmul_rn(my, my));
    const double cyz = __dsub_rn(__dmul_rn((double)syz, inv), __dmul_rn(my, mz));
    const double czz = __dsub_rn(__dmul_rn((double)szz, inv), __dmul_rn(mz, mz));
    const double tr = cxx + cyy + czz;
    double lam, nx, ny, nz;
    if constexpr (kExact) {
      smallest_eigen<double, 12>(cxx, cxy, cxz, cyy, cyz, czz, lam, nx, ny, nz);
      const double len = sqrt(nx * nx + ny * ny + nz * nz);
      nx /= len; ny /= len; nz /= len;
      lam = (tr != 0.0) ? fabs(lam / tr) : 0.0;
    } else {
      const float sc = (tr > 0.0) ? (float)(1.0 / tr) : 1.f;
      float fl, fx, fy, fz;
      smallest_eigen<float, 6>((float)cxx * sc, (float)cxy * sc, (float)cxz * sc, (float)cyy * sc, (float)cyz * sc,
                               (float)czz * sc, fl, fx, fy, fz);
      const float il = rsqrtf(fx * fx + fy * fy + fz * fz);
      nx = fx * il; ny = fy * il; nz = fz * il;
      lam = (tr > 0.0) ? fabsf(fl) : 0.f;
    }
    // flipNormalTowardsViewpoint: n.(vp - p) >= 0
    const double dot = nx * ((double)a.vpx - qx) + ny * ((double)a.vpy - qy) + nz * ((double)a.vpz - qz);
    if (dot < 0) { nx = -nx; ny = -ny; nz = -nz; }
    out = make_float4((float)nx, (float)ny, (float)nz, (float)lam);
  }
  if (pc.active) {
    a.nrm[pc.qi] = out;
    a.kcount[pc.qi] = k;
    if (a.nrm_in) {
      const int ii = g.perm[pc.qi];
      if (ii >= 0) a.nrm_in[ii] = out;
    }
  }
  unsigned long long ks = pc.active ? (unsigned long long)k : 0ull;
#pragma unroll
  for (int o = 16; o; o >>= 1) ks += __shfl_xor_sync(kFull, ks, o);
  if (lane == 0) {
    unsigned long long* slot = a.stats + 2 * (pid & (kStatSlots - 1));
    atomicAdd(slot, ks);
    atomicAdd(slot + 1, (unsigned long long)tested * (unsigned)pc.count);
  }
  }  // persistent packet loop
}

__global__ void fill_invalid_normals(float4* nrm, int* kcount, int begin, int end, const int* __restrict__ perm, float4* nrm_in) {
  int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  const float nan = __int_as_float(0x7fc00000);
  nrm[i] = make_float4(nan, nan, nan, nan);
  kcount[i] = 0;
  if (nrm_in && perm[i] >= 0) nrm_in[perm[i]] = make_float4(nan, nan, nan, nan);
}

}  // namespace

int run_normals(cab_ctx* ctx, float r, int max_nn, const float vp[3], const unsigned char* done, int hist_max_nn) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_normals: build the grid first");
  if (!(r > 0.f) || r > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_normals: radius %g exceeds the grid cell %g", (double)r, (double)ctx->cell);
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (int rc = reserve(ctx, ctx->b_nrm, (size_t)std::max(n, 1) * sizeof(float4))) return rc;
  if (int rc = reserve(ctx, ctx->b_kcount, (size_t)std::max(n, 1) * sizeof(int))) return rc;
  if (int rc = reserve(ctx, ctx->b_stats, kStatBytes)) return rc;
  const bool use_thr = max_nn > 0;
  if (use_thr)
    if (int rc = run_thresholds(ctx, r, max_nn, done, true)) return rc;
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_stats.p, 0, kStatBytes, st));
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
  NormalsArgs a{};
  a.done = done;
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  a.range = slab_packet_range(ctx, true);  // multi-GPU: own packets plus the rows around them, read on the device
  a.r = r;
  a.r2 = r * r;
  a.vpx = vp ? vp[0] : 0.f;
  a.vpy = vp ? vp[1] : 0.f;
  a.vpz = vp ? vp[2] : 0.f;
  a.one = 1.0f;
  a.nrm = (float4*)ctx->b_nrm.p;
  a.nrm_in = ctx->slab ? nullptr : ctx->fuse_nrm_in;
  a.kcount = (int*)ctx->b_kcount.p;
  a.thr_d2 = use_thr ? (const float*)ctx->b_thr_d2.p : nullptr;
  a.thr_idx = use_thr ? (const int*)ctx->b_thr_idx.p : nullptr;
  a.stats = (unsigned long long*)ctx->b_stats.p;
  // a slab's packet range lives on the device: launch the full persistent grid, surplus warps leave at once
  const int np = a.range ? std::max(1, (int)std::min<int64_t>(ctx->n_sorted, INT_MAX)) : a.p1 - a.p0;
  // The truncated fast RSD pass of the same call (cab_step_normals_rsd, cab_normals_rsd: one radius, normals untruncated,
  // RSD with max_nn) needs the d2 histogram of exactly these neighbourhoods: this traversal takes it along instead of a
  // traversal of its own (cab_topk.cu run_nn_hist then only settles the flagged packets).
  const bool with_hist = hist_max_nn > 0 && !use_thr && !ctx->cfg.exact && !ctx->slab && done == nullptr && np > 0;
  ctx->trunc_hist_valid = false;
  if (with_hist) {
    if (int rc = reserve(ctx, ctx->b_thr_idx, (size_t)std::max(n, 1) * sizeof(int))) return rc;
    if (int rc = reserve(ctx, ctx->b_thr_flag, 2 * (size_t)np + 32)) return rc;
    CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_thr_flag.p, 0, 2 * (size_t)np, st));
    a.code = (int*)ctx->b_thr_idx.p;
    a.flag = (unsigned char*)ctx->b_thr_flag.p;
    a.hist_max_nn = hist_max_nn;
    a.hist_scale = (float)kTruncBins / a.r2;
  }
  if (np > 0 && ctx->n_sorted > 0) {
    const dim3 blk(kWarpsPerBlock * kWarp);
    auto grid_for = [&](const void* fn) {
      int per_sm = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kWarpsPerBlock * kWarp, 0);
      return (unsigned)std::min<long long>((long long)std::max(per_sm, 1) * ctx->sm_count, (np + kWarpsPerBlock - 1) / kWarpsPerBlock);
    };
    if (ctx->cfg.exact && use_thr) normals_kernel<true, true><<<grid_for((const void*)normals_kernel<true, true>), blk, 0, st>>>(a);
    else if (ctx->cfg.exact) normals_kernel<true, false><<<grid_for((const void*)normals_kernel<true, false>), blk, 0, st>>>(a);
    else if (use_thr) normals_kernel<false, true><<<grid_for((const void*)normals_kernel<false, true>), blk, 0, st>>>(a);
    else if (with_hist) {
      const size_t hsmem = (size_t)kWarpsPerBlock * (kTruncBins + 1) * kWarp * sizeof(unsigned);
      CAB_CUDA(ctx, cudaFuncSetAttribute(normals_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
      int per_sm = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, normals_kernel<false, false, true>, kWarpsPerBlock * kWarp, hsmem);
      const unsigned grid = (unsigned)std::min<long long>((long long)std::max(per_sm, 1) * ctx->sm_count, (np + kWarpsPerBlock - 1) / kWarpsPerBlock);
      normals_kernel<false, false, true><<<grid, blk, hsmem, st>>>(a);
    }
    else normals_kernel<false, false><<<grid_for((const void*)normals_kernel<false, false>), blk, 0, st>>>(a);
    CAB_LAUNCH_CHECK(ctx);
  }
  if (!ctx->slab && n > ctx->n_valid) {  // non-finite points sit behind the sorted finite ones (a slab holds none)
    fill_invalid_normals<<<(n - ctx->n_valid + 255) / 256, 256, 0, st>>>((float4*)ctx->b_nrm.p, (int*)ctx->b_kcount.p,
                                                                        ctx->n_valid, n, (const int*)ctx->b_perm.p, a.nrm_in);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepStats0, ctx->b_stats.p, kStatBytes, cudaMemcpyDeviceToHost, st));
  ctx->have_normals = true;
  ctx->kcount_valid = !use_thr && done == nullptr;  // every query of this context's range counted, nothing truncated
  ctx->kcount_r = r;
  ctx->trunc_hist_valid = with_hist;
  ctx->trunc_hist_max_nn = hist_max_nn;
  if (ctx->defer_sync) return CAB_OK;
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return finish_pass_stats(ctx, 0);
}

int finish_pass_stats(cab_ctx* ctx, int pass) {
  if (pass == 0) CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.normals_ms, ctx->ev[2], ctx->ev[3]));
  else CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.rsd_ms, ctx->ev[4], ctx->ev[5]));
  read_stats(ctx, ctx->h_step + (pass == 0 ? kStepStats0 : kStepStats1));
  return CAB_OK;
}

}  // namespace cab
