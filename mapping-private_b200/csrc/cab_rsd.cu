// cab_rsd.cu -- Radius-based Surface Descriptor for every point of the cloud.
// Replaces HOT LOOP 1 + HOT LOOP 2 of cloud_algos::LocalRadiusEstimation::process
// (cloud_algos/src/radius_estimation.cpp:118-124 and :140-202): no neighbour list is ever
// materialised; each query keeps per-distance-bin extreme cosines in shared memory while the
// candidate chunks stream past, then solves the two one-parameter least-squares fits.
//
// Two kernels:
//  rsd_fast_kernel  fast mode (fp32, the default): every staged candidate is processed under a predicate -- packed
//           distance test and cosine, the bin row from two truncated products that bracket the reference's bin (the
//           exact thresholds only where they disagree), two shared-memory reductions into lane-private bins; the
//           truncated variant (max_nn) settles the target bin of the d2 histogram from a short per-query list.
//  rsd_kernel       exact mode (signed cosines, double acos) and the exact-threshold fallback of the truncation:
//    phase 1  per 32-candidate chunk: packed fp32x2 distance test -> one 32-bit hit mask per query
//    phase 2  only the hits are visited; the candidate's position and normal are fetched from the lane that staged it
//             with warp shuffles.
// Arithmetic notes:
//  * cosine is the reference's fp32 expression (nx*nx' + ny*ny') + nz*nz' (:153-155), not contracted.
//  * angle = acos(cosine) folded at pi/2 (:160-161) is monotone in |cosine|, so per bin only the
//    cosines of extreme |value| are tracked; acos is evaluated 2*ndiv times per query, not per pair.
//  * the distance bin floor(ndiv*(double)sqrtf(d2)/radius) (:165-168; `sqrt` of a float under `using namespace std;`
//    is std::sqrt(float)) is evaluated through exact
//    fp32 d2 thresholds computed on the host with the reference's double expression, clamped to
//    ndiv-1 (the reference indexes out of bounds at dist == radius).
//  * neighbours whose normal is not finite never update a bin (the reference's NaN comparisons are
//    all false, :158-172), so they are masked out when the chunk is staged.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "cab_internal.cuh"
#include "cab_traverse.cuh"

namespace cab {

namespace {

constexpr int kMaxDiv = 64;
constexpr float kBinBias = 1.0f / 1024.0f;  // >> error of the rsqrt bin estimate (<= 64 bins * 4e-7), << 1

struct RsdArgs {
  GridView g;
  int p0, p1;
  const int* range;        // optional device-side {p0, p1} (own packets of a slab); null: the values above
  const int* range_b;      // optional second device-side range, processed after the first (the two boundary layers)
  const SlabInfo* slab;    // with push.world > 0: where the rank's slice sits in the concatenated results
  PushTargets push;        // result exchange: every rank's copy of the concatenated arrays (peer memory)
  float r, r2;
  const float4* nrm;       // sorted order
  float2* out;             // sorted order (r_min, r_max)
  float* rdif;             // sorted order: (float)(max_radius - min_radius), the subtraction in double (:206)
  float* rmin_in;          // optional: r_min / r_max at the queries' input indices (cab_normals_rsd, input-order layout)
  float* rmax_in;
  const float* thr_d2;     // optional max_nn thresholds
  const int* thr_idx;
  const unsigned char* only;   // legacy kernel: optional per-packet flags (relative to p0), unflagged packets are skipped
  const unsigned char* skip;   // fast kernel: optional per-packet flags, flagged packets are left to the legacy kernel
  const int* trunc_code;       // truncated fast pass: per query target bin | kept << 8 (cab_topk.cu nn_hist_kernel)
  float trunc_scale;           // kTruncBins / r2
  const int* kcount;       // optional: in-radius neighbour counts of the normals pass at this radius (the statistics then need no count)
  const float* bin_thr;    // ndiv + 1 fp32 d2 thresholds, bin_thr[0] = -inf, bin_thr[ndiv] = +inf
  int ndiv;
  int flags;
  float bin_scale;         // ndiv / radius
  float magic;             // 2^23
  float bin_scale_lo, bin_scale_hi;  // the same biased low / high by 2^-18 relative (fast kernel: the two estimates bracket the bin)
  double radius, plane_radius;
  unsigned long long* stats;
  int work_slot;           // which packet work counter this launch pulls from (next_packet)
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

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f32x2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32x2(unsigned addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Result exchange fused into the kernel.  Replicated layout: the packet's normals, radii and input indices go straight
// into every rank's copy of the concatenated (sorted-order) result arrays -- 32 consecutive entries per store, over NVLink
// for the peers.  Input-range layout: every point's results go to ONE rank, the owner of its input index, already in
// input order (1/world of the traffic, no permutation left to undo).
__device__ __forceinline__ void push_results(const PushTargets& push, const SlabInfo* slab, const GridView& g,
                                             const PacketCtx& pc, const float4& nq, const float2 radii) {
  if (push.world <= 0 || !pc.active) return;
  const int input_index = g.perm[pc.qi];
  if (push.layout == CAB_COMM_LAYOUT_INPUT_RANGES) {
    // one owner per point: the rank whose input range holds it gets the results at the point's place in that range
    int p = (int)(((long long)input_index * push.world) / push.n);
    if (input_index >= push.lo[p + 1]) ++p;
    const int o = input_index - push.lo[p];
    push.nrm[p][o] = nq;
    push.rsd[p][o] = radii;
    return;
  }
  const int gpos = slab->gbase + (pc.qi - slab->q0);
  for (int p = 0; p < push.world; ++p) {
    push.nrm[p][gpos] = nq;
    push.rsd[p][gpos] = radii;
    push.perm[p][gpos] = input_index;
  }
}

template <bool kExact, bool kUseThr>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) rsd_kernel(const RsdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChunkTile* tiles = reinterpret_cast<ChunkTile*>(smem_raw);                 // [W]
  float2* bins = reinterpret_cast<float2*>(tiles + kWarpsPerBlock);          // [W][ndiv][32]
  float* thr = reinterpret_cast<float*>(bins + kWarpsPerBlock * a.ndiv * kWarp);  // [ndiv+1], padded to a multiple of 4
  int* self_slot = reinterpret_cast<int*>(thr + ((a.ndiv + 4) & ~3)) + (threadIdx.x & ~31);  // [W][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ndiv = a.ndiv;
  for (int i = threadIdx.x; i <= ndiv; i += blockDim.x) thr[i] = a.bin_thr[i];
  __shared__ double fbin[kMaxDiv];  // (di + 0.5) * radius / ndiv (:185), the same double expression, once per block
  if (threadIdx.x < ndiv) fbin[threadIdx.x] = (threadIdx.x + 0.5) * a.radius / ndiv;
  __syncthreads();
  const GridView& g = a.g;
  ChunkTile* tile = &tiles[warp];
  float2* my_b = bins + (size_t)warp * ndiv * kWarp + lane;  // bin b at my_b[b * 32]
  const unsigned bins_addr = smem_u32(my_b), thr_addr = smem_u32(thr);
  const float r2 = a.r2;
  const float bscale = a.bin_scale;
  const int last_bin = ndiv - 1;
  const int p0 = a.range ? a.range[0] : a.p0, n1 = (a.range ? a.range[1] : a.p1) - p0;
  const int pb = a.range_b ? a.range_b[0] : 0, n2 = a.range_b ? a.range_b[1] - pb : 0;
  for (;;) {
    const int w = next_packet(a.stats, lane, a.work_slot);
    if (w >= n1 + n2) break;
    const int pid = w < n1 ? p0 + w : pb + (w - n1);
    if (a.only && !a.only[pid - a.p0]) continue;
    const PacketCtx pc = load_packet(g, pid, lane, a.r, tile);
    const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
    const float4 nq = a.nrm[pc.qi];
    const bool q_ok = isfinite(nq.x) && isfinite(nq.y) && isfinite(nq.z);
    float td2 = INFINITY;
    int tidx = INT_MAX;
    if (kUseThr) {
      td2 = a.thr_d2[pc.qi];
      tidx = a.thr_idx[pc.qi];
    }
    // .x: cosine of smallest |value| (largest angle), .y: cosine of largest |value| (smallest angle)
    for (int b = 0; b < ndiv; ++b) my_b[b * kWarp] = make_float2(INFINITY, 0.f);
    if (a.flags & CAB_RSD_SEED_BIN0) my_b[0] = make_float2(1.f, 1.f);
    int k = 0;
    const int tested = for_each_chunk(g, pc, lane, tile, [&](int, const float4& c, int j, bool valid) {
      // this lane's own candidate: its normal (payload for the shuffles)
      const float4 cn = valid ? a.nrm[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      const unsigned finite_mask = __ballot_sync(kFull, isfinite(cn.x) && isfinite(cn.y) && isfinite(cn.z));
      unsigned mask = chunk_hit_mask(tile, qx, qy, qz, r2);
      if (!kUseThr) {
        k += __popc(mask);
        // non-finite normals never contribute
        mask &= q_ok ? finite_mask : 0u;
        // the query itself is skipped (:150 starts at ni = 1).  The packet's queries are the sorted points
        // [start, start + count): the lane that staged one of them tells the owner which bit to drop.
        self_slot[lane] = -1;
        __syncwarp();
        const int own = j - pc.start;
        if (valid && own >= 0 && own < pc.count) self_slot[own] = lane;
        __syncwarp();
        const int sb = self_slot[lane];
        if (sb >= 0) mask &= ~(1u << sb);
      }
      int iters = __reduce_max_sync(kFull, __popc(mask));
#pragma unroll 1
      for (; iters > 0; --iters) {
        const int m = 31 - __clz(mask);  // highest pending hit (-1 when none: shuffles read lane 31)
        const float cx = __shfl_sync(kFull, c.x, m), cy = __shfl_sync(kFull, c.y, m), cz = __shfl_sync(kFull, c.z, m);
        const float nx = __shfl_sync(kFull, cn.x, m), ny = __shfl_sync(kFull, cn.y, m), nz = __shfl_sync(kFull, cn.z, m);
        if (mask != 0) {
          mask ^= 1u << m;
          const float d2 = d2_rule(cx, cy, cz, qx, qy, qz);
          bool use = true;
          if (kUseThr) {
            // the query itself is skipped (:150 starts at ni = 1); only a zero distance can be the query
            use = !(d2 == 0.f && tile->idx[m] == pc.qi);
            const bool in = d2 < td2 || (d2 == td2 && g.perm[tile->idx[m]] <= tidx);
            k += in ? 1 : 0;
            use = use && in && q_ok && ((finite_mask >> m) & 1u);
          }
          if (use) {
            // clamping to [-1, 1] (:158-159) is monotone, so it is applied to the extremes only
            const float cs = __fadd_rn(__fadd_rn(__fmul_rn(nq.x, nx), __fmul_rn(nq.y, ny)), __fmul_rn(nq.z, nz));
            // distance bin: an estimate biased low by kBinBias (the rsqrt estimate is good to ~3e-5 bins), so the
            // bin is the estimate or the next one; the exact fp32 d2 threshold decides (d2 == 0 -> NaN -> 0)
            const int be = min(__float2int_rz(fmaf(d2 * rsqrt_approx(d2), bscale, -kBinBias)), last_bin);
            const int b = be + (d2 >= lds_f32(thr_addr + 4u * be + 4u) ? 1 : 0);
            const unsigned ba = bins_addr + (unsigned)(kWarp * sizeof(float2)) * b;
            if constexpr (kExact) {
              float2 v = lds_f32x2(ba);
              const float ac = fabsf(cs);
              if (ac < fabsf(v.x)) v.x = cs;
              if (ac >= fabsf(v.y)) v.y = cs;
              sts_f32x2(ba, v);
            } else {
              // the fast fit only needs |cosine| (acosf(|c|)): non-negative floats order like their bit
              // patterns, so the extremes are two fire-and-forget shared-memory atomics on the lane's own slot
              const unsigned ua = __float_as_uint(fabsf(cs));
              asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(ba), "r"(ua) : "memory");
              asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(ba + 4u), "r"(ua) : "memory");
            }
          }
        }
      }
    });

    // ---- least-squares fit of the min / max angle lines, radius_estimation.cpp:175-202 ----
    double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
    for (int di = 0; di < ndiv; ++di) {
      const float2 v = my_b[di * kWarp];
      if (fabsf(v.x) != INFINITY) {  // bin not empty (:181)
        const double p_min = fold_angle<kExact>(fminf(fmaxf(v.y, -1.f), 1.f));
        const double p_max = fold_angle<kExact>(fminf(fmaxf(v.x, -1.f), 1.f));
        const double f = fbin[di];
        Amint_Amin = __dadd_rn(Amint_Amin, __dmul_rn(p_min, p_min));
        Amint_d = __dadd_rn(Amint_d, __dmul_rn(p_min, f));
        Amaxt_Amax = __dadd_rn(Amaxt_Amax, __dmul_rn(p_max, p_max));
        Amaxt_d = __dadd_rn(Amaxt_d, __dmul_rn(p_max, f));
      }
    }
    const double max_radius = (Amint_Amin == 0) ? a.plane_radius : fmin(Amint_d / Amint_Amin, a.plane_radius);
    const double min_radius = (Amaxt_Amax == 0) ? a.plane_radius : fmin(Amaxt_d / Amaxt_Amax, a.plane_radius);
    float rmin = (float)min_radius, rmax = (float)max_radius;
    if (pc.active) a.rdif[pc.qi] = (float)(max_radius - min_radius);
    if (a.flags & CAB_RSD_SCALE_SORT) {
      const float x = rmax * 1.1f, y = rmin * 0.9f;
      rmin = fminf(x, y);
      rmax = fmaxf(x, y);
    }
    if (pc.active) a.out[pc.qi] = make_float2(rmin, rmax);
    if (a.rmin_in && pc.active) {
      const int ii = g.perm[pc.qi];
      if (ii >= 0) {
        a.rmin_in[ii] = rmin;
        a.rmax_in[ii] = rmax;
      }
    }
    push_results(a.push, a.slab, g, pc, nq, make_float2(rmin, rmax));
    unsigned long long ks = pc.active ? (unsigned long long)k : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) ks += __shfl_xor_sync(kFull, ks, o);
    if (lane == 0) {
      unsigned long long* slot = a.stats + 2 * (pid & (kStatSlots - 1));
      atomicAdd(slot, ks);
      atomicAdd(slot + 1, (unsigned long long)tested * (unsigned)pc.count);
    }
    __syncwarp();
  }
}

// ---- the fast-mode kernel (fp32, no max_nn): every staged candidate is processed under a predicate -------------------
// About 2 candidates in 5 are hits, and visiting only the hits (rsd_kernel above) pays for the compaction with a loop
// whose trip count is the largest hit count of the warp, six shuffles and a recomputed d2 per hit: ~36 issue slots per
// candidate tested.  Here the chunk's positions AND normals are staged in shared memory, and each candidate costs the
// distance test, the cosine, a three-instruction bin estimate (sqrt.approx, FFMA, F2I) and, under the hit predicate, one
// threshold load and two shared-memory reductions: ~22 slots, no divergence.  Differences from rsd_kernel, none in the
// results: chunks are consecutive stream positions (coalesced loads; the packet's own queries then sit in at most two
// chunks, and only those run the variant of the loop that knows the self bit); candidates whose normal is not finite
// are staged far away (they never contribute, :158-172) and counted in a rare side path; the bins are two arrays
// [bin][lane] (min |cos|, max |cos|), so that a warp's reductions never meet in a bank whatever bins its lanes hit.
constexpr int kFastThr = 260;  // entries of the fast kernel's threshold table (257 used; the truncated variant reads it)
// One bin row of a warp in the fast kernel: min |cos| [32], max |cos| [32] (the reductions below reach the second half
// with a +128 byte offset)
constexpr int kRowWords = 2 * kWarp;
constexpr unsigned kRowBytes = kRowWords * sizeof(unsigned);
struct alignas(16) FastTile {
  float x[kWarp], y[kWarp], z[kWarp];
  float nx[kWarp], ny[kWarp], nz[kWarp];
  int idx[kWarp];  // sorted index of the staged candidate (truncated pass: the target bin's candidates are listed by it)
  int self_slot[kWarp];
  int run_begin[12];
  int run_cum[12];
};

// Truncated pass (kTrunc).  A candidate's place in the 64-bin histogram of nn_hist_kernel is min(int(d2 * trunc_scale), 63);
// the query's target bin t1 holds neighbour number max_nn + 1.  Both edges of that bin are turned into exact fp32 d2
// cut-offs once per query (bin_edge below: the smallest d2 whose product reaches the bin -- the comparison against it
// decides exactly what the product would), so that the loop is the untruncated loop with the radius test d2 <= r2 replaced
// by d2 < cut (a neighbour below the target bin is kept outright), plus one bit per candidate of the target bin
// (cut <= d2 < hi) in `tmask`: the caller lists those few candidates after the chunk and settles them after the
// traversal.  Beyond the bin a candidate is dropped.  A query that keeps everything has cut = hi = nextafter(r2).
// The ambiguous candidates of fast_chunk (below): the low estimate n = lo - 2^23's bits is the reference's bin or the one
// below it; thr[b] is the smallest fp32 d2 of bin b (thr[ndiv] = +inf: a neighbour never leaves the last bin).  What the
// unambiguous path decided for the candidate otherwise -- dropped by the truncation, the query itself, beyond the radius
// by more than the margin: its row is the spare one already -- stands.
__device__ __noinline__ unsigned exact_row(unsigned lo, unsigned vb, float d2, float r2, const float* thr, int ndiv, unsigned row_k,
                                           unsigned addr) {
  const int n = (int)(lo - 0x4B000000u);
  if (n >= ndiv || vb == 0x4B000000u + (unsigned)ndiv) return addr;
  const int b = n + (d2 >= thr[n + 1] ? 1 : 0);
  return (0x4B000000u + (unsigned)(d2 <= r2 ? min(b, ndiv - 1) : ndiv)) * kRowBytes + row_k;
}

template <bool kSelf, bool kCount, bool kTrunc, int kUnroll>
__device__ __forceinline__ int fast_chunk(const FastTile* tile, float qx, float qy, float qz, float nqx, float nqy, float nqz,
                                          float r2, float s_lo, float s_hi, float magic, float vmax, unsigned row_k,
                                          const float* thr, int ndiv, int sb, float hi, unsigned& tmask) {
  const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
  const f32x2 nqx2 = pack2(nqx, nqx), nqy2 = pack2(nqy, nqy), nqz2 = pack2(nqz, nqz);
  const float4* tx = reinterpret_cast<const float4*>(tile->x);
  const float4* ty = reinterpret_cast<const float4*>(tile->y);
  const float4* tz = reinterpret_cast<const float4*>(tile->z);
  const float4* tnx = reinterpret_cast<const float4*>(tile->nx);
  const float4* tny = reinterpret_cast<const float4*>(tile->ny);
  const float4* tnz = reinterpret_cast<const float4*>(tile->nz);
  int k = 0;
  unsigned tm = 0;  // kTrunc: this chunk's candidates of the target bin
  // One group of four candidates: everything up to the bin address is plain arithmetic on independent chains (the
  // compiler interleaves them); only the two reductions and the count sit under the hit predicate.  No "memory" clobber
  // on the reductions: the warp barriers around the chunk order them against the plain accesses of the bins.
  // (the truncated variant is unrolled half as far: its body is longer, and fully unrolled the kernel's hot code no longer
  // fits the instruction cache -- ncu showed 3 issue slots in 10 waiting for instructions)
  const f32x2 s2 = pack2(s_lo, s_hi), magic2 = pack2(magic, magic);
#pragma unroll(kUnroll)
  for (int g4 = 0; g4 < kWarp / 4; ++g4) {
    const float4 X = tx[g4], Y = ty[g4], Z = tz[g4];
    const float4 NX = tnx[g4], NY = tny[g4], NZ = tnz[g4];
    float d2[4];
    {
      const f32x2 dx = sub2(pack2(X.x, X.y), qx2), dy = sub2(pack2(Y.x, Y.y), qy2), dz = sub2(pack2(Z.x, Z.y), qz2);
      unpack2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), d2[0], d2[1]);
    }
    {
      const f32x2 dx = sub2(pack2(X.z, X.w), qx2), dy = sub2(pack2(Y.z, Y.w), qy2), dz = sub2(pack2(Z.z, Z.w), qz2);
      unpack2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), d2[2], d2[3]);
    }
    // radius_estimation.cpp:153-155, the fp32 expression as written (products and sums rounded separately), two
    // candidates per packed instruction
    float cs[4];
    {
      const f32x2 a = mul2(nqx2, pack2(NX.x, NX.y)), b = mul2(nqy2, pack2(NY.x, NY.y)), c = mul2(nqz2, pack2(NZ.x, NZ.y));
      unpack2(add2(add2(a, b), c), cs[0], cs[1]);
    }
    {
      const f32x2 a = mul2(nqx2, pack2(NX.z, NX.w)), b = mul2(nqy2, pack2(NY.z, NY.w)), c = mul2(nqz2, pack2(NZ.z, NZ.w));
      unpack2(add2(add2(a, b), c), cs[2], cs[3]);
    }
    // The candidates' bin rows, straight out of the floating-point pipe: v = 2^23 + floor(root * scale) (fma.rz on the
    // magic constant: the integer sits in the low mantissa bits), once with the scale ndiv / radius biased low and once
    // biased high by 2^-18 relative -- far more than the errors of sqrt.approx (2^-22), of the scale's rounding and of the
    // reference's own (double)sqrtf(d2) (2^-24) together.  Where the two floors agree, every value in between has that
    // floor: it IS the reference's bin floor(ndiv * (double)sqrtf(d2) / radius) (:165-168), no threshold needed.  They
    // disagree for about one candidate in 10^4 (d2 within 2^-17 relative of a bin edge, the radius included): the warp
    // votes once per group of four and only then consults the exact fp32 d2 thresholds (computed on the host with the
    // reference's double expression).  ONE integer multiply-add turns the float's bits into the lane's address of the
    // row (row_k holds the lane's column minus 2^23's bits times the stride, modulo 2^32); misses and staged-away
    // candidates (d2 = inf) are clamped to the spare row.
    // (ptxas turns a predicated shared-memory reduction into a branch around it, so the reductions below are
    // unconditional and the misses go to the spare row.)
    unsigned ua[4], addr[4], lo[4], vb[4];
    unsigned amb = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ua[i] = __float_as_uint(cs[i]) & 0x7fffffffu;  // |cos| on the integer pipe (the FP32 pipes are the busier ones)
      float root;
      asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(root) : "f"(d2[i]));
      float vlo, vhi;  // both estimates from one packed fma.rz
      {
        f32x2 v2;
        asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(v2) : "l"(pack2(root, root)), "l"(s2), "l"(magic2));
        unpack2(v2, vlo, vhi);
      }
      lo[i] = __float_as_uint(vlo);
      amb |= lo[i] ^ __float_as_uint(vhi);
      float v = fminf(vlo, vmax);
      if (kTrunc) {
        // r2 holds the query's cut: "d2 < cut" keeps the candidate (it lies inside the radius); a dropped one goes to the
        // spare row; cut <= d2 < hi marks it in the chunk's target mask
        asm("{\n\t.reg .pred q;\n\t"
            "setp.ge.f32 q, %1, %2;\n\t"
            "setp.lt.and.f32 q, %1, %3, q;\n\t"
            "@q or.b32 %0, %0, %4;\n\t}"
            : "+r"(tm)
            : "f"(d2[i]), "f"(r2), "f"(hi), "r"(1u << (4 * g4 + i)));
        asm("{\n\t.reg .pred p, s;\n\t"
            "setp.ne.s32 s, %4, %5;\n\t"
            "setp.lt.and.f32 p, %2, %3, s;\n\t"
            "@!p mov.f32 %0, %6;\n\t"
            "@p add.s32 %1, %1, 1;\n\t}"
            : "+f"(v), "+r"(k)
            : "f"(d2[i]), "f"(r2), "r"(4 * g4 + i), "r"(sb), "f"(vmax));
      } else if (kSelf) {
        asm("{\n\t.reg .pred p, s;\n\t"
            "setp.eq.s32 s, %4, %5;\n\t"
            "@s mov.f32 %0, %6;\n\t"
            "setp.le.and.f32 p, %2, %3, !s;\n\t"
            "@p add.s32 %1, %1, 1;\n\t}"
            : "+f"(v), "+r"(k)
            : "f"(d2[i]), "f"(r2), "r"(4 * g4 + i), "r"(sb), "f"(vmax));
      } else if (kCount) {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.le.f32 p, %1, %2;\n\t"
            "@p add.s32 %0, %0, 1;\n\t}"
            : "+r"(k)
            : "f"(d2[i]), "f"(r2));
      }
      vb[i] = __float_as_uint(v);
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr[i]) : "r"(vb[i]), "n"(kRowBytes), "r"(row_k));
    }
    if (__any_sync(kFull, amb != 0)) {
      // some lane's candidate sits within 2^-17 of a bin edge: its row from the exact thresholds (out of line: rare)
#pragma unroll
      for (int i = 0; i < 4; ++i) addr[i] = exact_row(lo[i], vb[i], d2[i], kTrunc ? INFINITY : r2, thr, ndiv, row_k, addr[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(addr[i]), "r"(ua[i]));
      asm volatile("red.shared.max.u32 [%0+128], %1;" ::"r"(addr[i]), "r"(ua[i]));
    }
  }
  if (kTrunc) tmask = tm;
  return k;
}

// Truncated pass, a chunk that holds a candidate without a finite normal (rare: a point with fewer than three
// neighbours).  Such a candidate is a neighbour like any other -- it counts towards max_nn and may sit in the target bin
// -- but forms no pair (radius_estimation.cpp:158-172 never sees a normal it could use).  The predicated loop above has no
// room for that rule, so the chunk is walked candidate by candidate with the rules spelled out; same bins, same list.
__device__ __forceinline__ int trunc_slow_chunk(const FastTile* tile, float qx, float qy, float qz, float nqx, float nqy, float nqz,
                                                 float cut, float hi, const float* thr, int ndiv, unsigned* my_min, unsigned* my_max,
                                                 int sb, unsigned& tmask) {
  int k = 0;
  unsigned tm = 0;
  for (int m = 0; m < kWarp; ++m) {
    const float d2 = d2_rule(tile->x[m], tile->y[m], tile->z[m], qx, qy, qz);
    if (!(d2 < cut)) {
      if (d2 < hi) tm |= 1u << m;  // the target bin: listed by the caller, the query itself included
      continue;
    }
    if (m == sb) continue;  // the query itself: counted by the caller
    ++k;
    const float nx = tile->nx[m], ny = tile->ny[m], nz = tile->nz[m];
    if (!(isfinite(nx) && isfinite(ny) && isfinite(nz))) continue;
    const float cs = __fadd_rn(__fadd_rn(__fmul_rn(nqx, nx), __fmul_rn(nqy, ny)), __fmul_rn(nqz, nz));
    const unsigned ua = __float_as_uint(fabsf(cs));
    int b = 0;
    while (b < ndiv - 1 && d2 >= thr[b + 1]) ++b;
    my_min[b * kRowWords] = min(my_min[b * kRowWords], ua);
    my_max[b * kRowWords] = max(my_max[b * kRowWords], ua);
  }
  tmask = tm;
  return k;
}

// smallest fp32 d2 >= 0 whose product with `scale` (rounded to nearest, as the histogram computes it) reaches t
__device__ __forceinline__ float bin_edge(float t, float scale) {
  if (!(t > 0.f)) return 0.f;
  float x = __fdiv_rn(t, scale);
  while (x > 0.f && __fmul_rn(x, scale) >= t) x = __uint_as_float(__float_as_uint(x) - 1u);
  while (__fmul_rn(x, scale) < t) x = __uint_as_float(__float_as_uint(x) + 1u);
  return x;
}

template <bool kCount, bool kTrunc, int kUnroll>
__global__ void __launch_bounds__(kWarpsPerBlock * kWarp) rsd_fast_kernel(const RsdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastTile* tiles = reinterpret_cast<FastTile*>(smem_raw);                          // [W]
  unsigned* bins = reinterpret_cast<unsigned*>(tiles + kWarpsPerBlock);                    // [W][ndiv + 1][2][32], row ndiv: the misses
  float* thr = reinterpret_cast<float*>(bins + kWarpsPerBlock * kRowWords * (a.ndiv + 1));  // [257], +inf from ndiv on
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ndiv = a.ndiv;
  // kTrunc: [W][kTruncCap][32], a lane's list of its target bin's candidates: their positions in the packet's candidate
  // stream (16 bits: the histogram kernels leave packets with 65536 candidates or more to the exact-threshold path)
  unsigned short* lists = reinterpret_cast<unsigned short*>(thr + kFastThr);
  for (int i = threadIdx.x; i < kFastThr; i += blockDim.x) thr[i] = i <= ndiv ? a.bin_thr[i] : INFINITY;
  // the bins' distances (di + 0.5) * radius / ndiv (:185): the same double expression, evaluated once per block instead of
  // once per bin and query
  __shared__ double fbin[kMaxDiv];
  if (threadIdx.x < ndiv) fbin[threadIdx.x] = (threadIdx.x + 0.5) * a.radius / ndiv;
  __syncthreads();
  const GridView& g = a.g;
  FastTile* tile = &tiles[warp];
  unsigned short* lst = lists + (size_t)warp * kTruncCap * kWarp + lane;
  unsigned* my_min = bins + (size_t)warp * kRowWords * (ndiv + 1) + lane;  // bin b: min |cos| at my_min[b * 64], max at my_max[b * 64]
  unsigned* my_max = my_min + kWarp;
  const unsigned bins_addr = smem_u32(my_min);
  const float r2 = a.r2;
  // (2^23 arrives as a kernel argument: as an immediate it would take the operand slot the scales' constant-bank
  // addresses need, and every group of candidates would load the scales again)
  const float s_lo = a.bin_scale_lo, s_hi = a.bin_scale_hi, magic = a.magic, vmax = 8388608.f + (float)ndiv;
  const unsigned row_k = bins_addr - 0x4B000000u * kRowBytes;  // modulo 2^32: the multiply-add of fast_chunk wraps the same way
  const int p0 = a.range ? a.range[0] : a.p0, n1 = (a.range ? a.range[1] : a.p1) - p0;
  const int pb = a.range_b ? a.range_b[0] : 0, n2 = a.range_b ? a.range_b[1] - pb : 0;
  for (;;) {
    const int w = next_packet(a.stats, lane, a.work_slot);
    if (w >= n1 + n2) break;
    const int pid = w < n1 ? p0 + w : pb + (w - n1);
    if (kTrunc && a.skip && a.skip[pid - a.p0]) continue;  // left to the exact-threshold path
    // the packet and its candidate runs (load_packet works on a ChunkTile; the run tables sit at the same place here)
    PacketCtx pc;
    float cut = 0.f, hi = 0.f;  // kTrunc
    int need = 0;
    {
      const Packet pk = g.packets[pid];
      const Domain dm = g.domains[pk.domain];
      pc.start = pk.start;
      pc.count = pk.count;
      pc.active = lane < pk.count;
      pc.qi = pk.start + min(lane, pk.count - 1);
      pc.q = g.pos[pc.qi];
      const float xmin = warp_min(pc.q.x), xmax = warp_max(pc.q.x);
      float rc = a.r * 1.00001f;
      if (kTrunc) {
        // the query's target bin of the d2 histogram as two exact d2 cut-offs (see fast_chunk), and how many of that
        // bin's candidates are kept
        const int code = a.trunc_code[pc.qi];
        const int t1 = code & 255;
        need = code >> 8;
        const float r2n = __uint_as_float(__float_as_uint(a.r2) + 1u);  // d2 < r2n is d2 <= r2
        if (t1 == 255) {
          cut = hi = r2n;
        } else {
          cut = bin_edge((float)t1, a.trunc_scale);
          hi = t1 == kTruncBins - 1 ? r2n : fminf(bin_edge((float)(t1 + 1), a.trunc_scale), r2n);
        }
        // no query of the packet keeps a neighbour at or beyond `hi`: the candidate runs are trimmed to the largest such
        // distance instead of the radius
        rc = fminf(rc, sqrtf(warp_max(hi)) * 1.00001f);
      }
      const int cy = pk.row_local % dm.ny, cz = pk.row_local / dm.ny;
      const int cxlo = max((xfine_coord(xmin, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift) - 1, 0);
      const int cxhi = min((xfine_coord(xmax, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide) >> dm.xshift) + 1, dm.nx - 1);
      const int xf_lo = xfine_coord(xmin - rc, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
      const int xf_hi = xfine_coord(xmax + rc, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
      const int t = lane & 15;
      int lo = 0, hi = 0;
      if (t < 9) {
        const int y = cy + t % 3 - 1, z = cz + t / 3 - 1;
        if (row_in_table(dm, y, z)) {
          const long long c = dm.cell_base + ((long long)z * dm.ny + y) * dm.nx;
          lo = g.cell_start[c + cxlo];
          hi = g.cell_start[c + cxhi + 1];
        }
      }
      const int key = lane < 16 ? xf_lo : xf_hi + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int xf = xfine_coord(g.pos[mid].x, dm.ox, g.inv_cell, dm.nx, dm.xshift, dm.xwide);
        if (xf < key) lo = mid + 1; else hi = mid;
      }
      const int end = __shfl_sync(kFull, lo, (lane + 16) & 31);
      const int len = (lane < 9) ? max(end - lo, 0) : 0;
      int cum = len;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int v = __shfl_up_sync(kFull, cum, o);
        if (lane >= o) cum += v;
      }
      __syncwarp();
      if (lane < 9) {
        tile->run_begin[lane] = lo;
        tile->run_cum[lane] = cum - len;
      }
      pc.total = __shfl_sync(kFull, cum, 8);
      if (lane == 9) tile->run_cum[9] = pc.total;
      __syncwarp();
    }
    const float qx = pc.q.x, qy = pc.q.y, qz = pc.q.z;
    const float4 nq = a.nrm[pc.qi];
    const bool q_ok = isfinite(nq.x) && isfinite(nq.y) && isfinite(nq.z);
    for (int b = 0; b < ndiv; ++b) {
      my_min[b * kRowWords] = 0x7f800000u;  // +inf: empty
      my_max[b * kRowWords] = 0u;
    }
    if (a.flags & CAB_RSD_SEED_BIN0) {
      my_min[0] = 0x3f800000u;
      my_max[0] = 0x3f800000u;
    }
    int k = 0;
    int lcnt = 0;  // truncated pass: candidates of the target bin listed so far
    const int nchunks = (pc.total + kWarp - 1) / kWarp;
#pragma unroll 1
    for (int c0 = 0; c0 < nchunks; ++c0) {
      const int p = c0 * kWarp + lane;
      const bool valid = p < pc.total;
      int t = (p >= tile->run_cum[4]) ? 4 : 0;
      t += (p >= tile->run_cum[t + 2]) ? 2 : 0;
      t += (p >= tile->run_cum[t + 1]) ? 1 : 0;
      t += (t == 7 && p >= tile->run_cum[8]) ? 1 : 0;
      const int j = valid ? tile->run_begin[t] + (p - tile->run_cum[t]) : -1;
      float4 c = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 0.f), cn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        c = g.pos[j];
        cn = a.nrm[j];
      }
      const bool finite_n = isfinite(cn.x) && isfinite(cn.y) && isfinite(cn.z);
      const int own = j - pc.start;
      const bool is_own = valid && own >= 0 && own < pc.count;
      __syncwarp();
      // a candidate without a normal never contributes: staged far away -- except in the truncated pass, where it still
      // counts towards max_nn (its chunk takes the spelled-out path below)
      tile->x[lane] = (valid && (finite_n || kTrunc)) ? c.x : 3.0e30f;
      tile->y[lane] = c.y;
      tile->z[lane] = c.z;
      tile->nx[lane] = cn.x;
      tile->ny[lane] = cn.y;
      tile->nz[lane] = cn.z;
      if (kTrunc) tile->idx[lane] = j;
      const unsigned own_mask = __ballot_sync(kFull, is_own);
      const unsigned odd_mask = __ballot_sync(kFull, valid && !finite_n);
      if (own_mask) {
        tile->self_slot[lane] = -1;
        __syncwarp();
        if (is_own) tile->self_slot[own] = lane;
      }
      __syncwarp();
      if (kTrunc) {
        const int sb = own_mask ? tile->self_slot[lane] : -1;
        unsigned tmask = 0;
        if (odd_mask)
          k += trunc_slow_chunk(tile, qx, qy, qz, nq.x, nq.y, nq.z, cut, hi, thr, ndiv, my_min, my_max, sb, tmask);
        else  // one body for chunks with and without queries of the packet (sb = -1 matches no slot): half the code
          k += fast_chunk<true, true, true, kUnroll>(tile, qx, qy, qz, nq.x, nq.y, nq.z, cut, s_lo, s_hi, magic, vmax, row_k, thr, ndiv, sb, hi, tmask);
        k += (sb >= 0 && cut > 0.f) ? 1 : 0;  // the query itself, unless it sits in the target bin (then the list has it)
        while (tmask) {  // the chunk's candidates of the target bin (a handful per query and traversal): listed for the settlement
          const int m = __ffs(tmask) - 1;
          tmask &= tmask - 1;
          if (lcnt < kTruncCap) lst[lcnt * kWarp] = (unsigned short)(c0 * kWarp + m);
          ++lcnt;
        }
      } else if (own_mask) {
        unsigned dummy_mask;
        k += fast_chunk<true, true, false, kUnroll>(tile, qx, qy, qz, nq.x, nq.y, nq.z, r2, s_lo, s_hi, magic, vmax, row_k, thr, ndiv, tile->self_slot[lane], 0.f, dummy_mask);
        k += tile->self_slot[lane] >= 0 ? 1 : 0;  // the query itself is a neighbour of the radius search (:120), just not a pair (:150)
      } else {
        unsigned dummy_mask;
        k += fast_chunk<false, kCount, false, kUnroll>(tile, qx, qy, qz, nq.x, nq.y, nq.z, r2, s_lo, s_hi, magic, vmax, row_k, thr, ndiv, -1, 0.f, dummy_mask);
      }
      if (!kTrunc && odd_mask) {  // rare: neighbours without a normal still count as neighbours
        unsigned mm = odd_mask;
        while (mm) {
          const int m = __ffs(mm) - 1;
          mm &= mm - 1;
          const float cx = __shfl_sync(kFull, c.x, m), cy = __shfl_sync(kFull, c.y, m), cz = __shfl_sync(kFull, c.z, m);
          const int cj = __shfl_sync(kFull, j, m);
          if (d2_rule(cx, cy, cz, qx, qy, qz) <= r2 && cj != pc.qi) k += 1;  // the query itself is counted above
        }
      }
    }

    __syncwarp();  // the reductions above are complete before the bins are read back
    if (kTrunc && lcnt > 0) {
      // The target bin: keep the `need` smallest (d2, input index) of its candidates (the rule of the radius search with
      // max_nn, radius_estimation.cpp:120); the query itself counts among them but forms no pair (:150).
      const int L = min(lcnt, kTruncCap);
      // the candidates' d2 (the bits the traversal saw): recomputed once each into the warp's tile, whose eight rows the
      // traversal no longer needs; entries beyond the eighth (rare) are recomputed where they are used
      float* scratch = reinterpret_cast<float*>(tile) + lane;
      // sorted index of list entry e: its stream position looked up in the packet's run table (which sits behind the
      // tile's eight rows and is still in place)
      auto idx_of = [&](int e) {
        const int p = lst[e * kWarp];
        int t = (p >= tile->run_cum[4]) ? 4 : 0;
        t += (p >= tile->run_cum[t + 2]) ? 2 : 0;
        t += (p >= tile->run_cum[t + 1]) ? 1 : 0;
        t += (t == 7 && p >= tile->run_cum[8]) ? 1 : 0;
        return tile->run_begin[t] + (p - tile->run_cum[t]);
      };
      auto d2_of = [&](int e) {
        if (e < 8) return scratch[e * kWarp];
        const float4 pe = g.pos[idx_of(e)];
        return d2_rule(pe.x, pe.y, pe.z, qx, qy, qz);
      };
      for (int e = 0; e < min(L, 8); ++e) {
        const float4 pe = g.pos[idx_of(e)];
        scratch[e * kWarp] = d2_rule(pe.x, pe.y, pe.z, qx, qy, qz);
      }
      for (int e = 0; e < L; ++e) {
        const int je = idx_of(e);
        const float de = d2_of(e);
        int rank = 0;
        for (int f = 0; f < L; ++f) {
          const float df = d2_of(f);
          if (df < de) ++rank;
          else if (df == de && f != e && g.perm[idx_of(f)] < g.perm[je]) ++rank;
        }
        if (rank >= need) continue;
        ++k;
        if (je == pc.qi) continue;
        const float4 cn = a.nrm[je];
        if (!(isfinite(cn.x) && isfinite(cn.y) && isfinite(cn.z))) continue;  // a neighbour, but no pair
        const float cs = __fadd_rn(__fadd_rn(__fmul_rn(nq.x, cn.x), __fmul_rn(nq.y, cn.y)), __fmul_rn(nq.z, cn.z));
        const unsigned ua = __float_as_uint(fabsf(cs));
        int b = 0;
        while (b < ndiv - 1 && de >= thr[b + 1]) ++b;
        my_min[b * kRowWords] = min(my_min[b * kRowWords], ua);
        my_max[b * kRowWords] = max(my_max[b * kRowWords], ua);
      }
    }
    // ---- least-squares fit of the min / max angle lines, radius_estimation.cpp:175-202 ----
    double Amint_Amin = 0, Amint_d = 0, Amaxt_Amax = 0, Amaxt_d = 0;
    if (q_ok) {
      for (int di = 0; di < ndiv; ++di) {
        const float lo = __uint_as_float(my_min[di * kRowWords]), hi = __uint_as_float(my_max[di * kRowWords]);
        if (lo != INFINITY) {  // bin not empty (:181)
          const double p_min = fold_angle<false>(fminf(hi, 1.f));
          const double p_max = fold_angle<false>(fminf(lo, 1.f));
          const double f = fbin[di];
          Amint_Amin = __dadd_rn(Amint_Amin, __dmul_rn(p_min, p_min));
          Amint_d = __dadd_rn(Amint_d, __dmul_rn(p_min, f));
          Amaxt_Amax = __dadd_rn(Amaxt_Amax, __dmul_rn(p_max, p_max));
          Amaxt_d = __dadd_rn(Amaxt_d, __dmul_rn(p_max, f));
        }
      }
    }
    const double max_radius = (Amint_Amin == 0) ? a.plane_radius : fmin(Amint_d / Amint_Amin, a.plane_radius);
    const double min_radius = (Amaxt_Amax == 0) ? a.plane_radius : fmin(Amaxt_d / Amaxt_Amax, a.plane_radius);
    float rmin = (float)min_radius, rmax = (float)max_radius;
    if (pc.active) a.rdif[pc.qi] = (float)(max_radius - min_radius);
    if (a.flags & CAB_RSD_SCALE_SORT) {
      const float x = rmax * 1.1f, y = rmin * 0.9f;
      rmin = fminf(x, y);
      rmax = fmaxf(x, y);
    }
    if (pc.active) a.out[pc.qi] = make_float2(rmin, rmax);
    if (a.rmin_in && pc.active) {
      const int ii = g.perm[pc.qi];
      if (ii >= 0) {
        a.rmin_in[ii] = rmin;
        a.rmax_in[ii] = rmax;
      }
    }
    push_results(a.push, a.slab, g, pc, nq, make_float2(rmin, rmax));
    if (!kCount) k = a.kcount[pc.qi];  // same radius, same rule, no truncation: the normals pass counted these neighbours
    unsigned long long ks = pc.active ? (unsigned long long)k : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) ks += __shfl_xor_sync(kFull, ks, o);
    if (lane == 0) {
      unsigned long long* slot = a.stats + 2 * (pid & (kStatSlots - 1));
      atomicAdd(slot, ks);
      atomicAdd(slot + 1, (unsigned long long)pc.total * (unsigned)pc.count);
    }
    __syncwarp();
  }
}

// CAB_STEP_INPUT_ORDER: the points that are nobody's query (non-finite, sorted behind the finite ones) in input order
__global__ void fill_invalid_input_order(const int* __restrict__ perm, int begin, int end, float v, float4* __restrict__ nrm,
                                         float2* __restrict__ rsd) {
  const int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  const int j = perm[i];
  const float nan = __int_as_float(0x7fc00000);
  nrm[j] = make_float4(nan, nan, nan, nan);
  rsd[j] = make_float2(v, v);
}

__global__ void fill_invalid_rsd(float2* out, float* rdif, int begin, int end, float v, const int* __restrict__ perm, float* rmin_in,
                                 float* rmax_in) {
  int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < end) {
    out[i] = make_float2(v, v);
    rdif[i] = 0.f;
    if (rmin_in) {
      const int j = perm[i];
      if (j >= 0) {
        rmin_in[j] = v;
        rmax_in[j] = v;
      }
    }
  }
}

// Smallest fp32 d2 whose reference bin floor(ndiv*(double)sqrtf(d2)/radius) is >= b.
float bin_threshold(int b, int ndiv, double radius, float r2) {
  auto bin_of = [&](float d2) { return (int)std::floor(ndiv * (double)std::sqrt(d2) / radius); };
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

template <bool kExact, bool kUseThr>
int launch_rsd(cab_ctx* ctx, const RsdArgs& a, unsigned blocks, size_t smem, cudaStream_t ks) {
  CAB_CUDA(ctx, cudaFuncSetAttribute(rsd_kernel<kExact, kUseThr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CAB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rsd_kernel<kExact, kUseThr>, kWarpsPerBlock * kWarp, smem));
  blocks = std::min<unsigned>(blocks, (unsigned)std::max(per_sm, 1) * ctx->sm_count);  // persistent warps
  rsd_kernel<kExact, kUseThr><<<blocks, kWarpsPerBlock * kWarp, smem, ks>>>(a);
  CAB_LAUNCH_CHECK(ctx);
  return CAB_OK;
}

}  // namespace

int run_rsd(cab_ctx* ctx, double r, int max_nn, int ndiv, double plane_radius, int flags, int phase) {
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_rsd: build the grid first");
  if (!ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "cab_rsd: missing normals");
  const float rf = (float)r;
  if (!(rf > 0.f) || rf > ctx->cell * 1.0000001f)
    return fail(ctx, CAB_ERR_ARG, "cab_rsd: radius %g exceeds the grid cell %g", r, (double)ctx->cell);
  if (ndiv < 1 || ndiv > kMaxDiv) return fail(ctx, CAB_ERR_ARG, "cab_rsd: distance_div must be in [1, %d]", kMaxDiv);
  const int n = (int)ctx->n;
  // phase 2 (the boundary packets of a split pass) is launched on the copy stream with its own work counter: it runs
  // beside the tail of phase 1 instead of after it
  cudaStream_t st = ctx->stream;
  static const bool serial_phases = std::getenv("CAB_SERIAL_PHASES") != nullptr;  // A/B switch
  cudaStream_t ks = phase == 2 && !serial_phases ? ctx->copy_stream : ctx->stream;
  const int slot = phase == 2 ? 1 : 0;
  if (int rc = reserve(ctx, ctx->b_rsd, (size_t)std::max(n, 1) * sizeof(float2))) return rc;
  if (int rc = reserve(ctx, ctx->b_rdif, (size_t)std::max(n, 1) * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_stats, kStatBytes)) return rc;
  const bool use_thr = max_nn > 0;
  static const bool legacy = std::getenv("CAB_RSD_LEGACY") != nullptr;  // A/B switch for profiling
  // fast mode with max_nn: one histogram traversal, then the RSD traversal settles the truncation itself
  const bool trunc_fast = use_thr && !ctx->cfg.exact && !legacy;
  if (phase == 2) {
    // second launch of a split pass: everything was set up by phase 1 (both work counters included)
  } else if (trunc_fast) {
    if (int rc = run_nn_hist(ctx, rf, max_nn)) return rc;
  } else if (use_thr) {
    if (int rc = run_thresholds(ctx, rf, max_nn)) return rc;
  }
  // bin thresholds
  float thr[kMaxDiv + 1];
  const float r2 = rf * rf;
  thr[0] = -INFINITY;
  for (int b = 1; b < ndiv; ++b) thr[b] = bin_threshold(b, ndiv, r, r2);
  thr[ndiv] = INFINITY;
  // staged through its own device / pinned slots: in a deferred step nothing before it in the stream has been waited for
  if (int rc = reserve(ctx, ctx->b_stats2, sizeof(thr))) return rc;
  if (phase != 2) {
    std::memcpy(ctx->h_step + kStepThr, thr, sizeof(float) * (ndiv + 1));
    CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_stats2.p, ctx->h_step + kStepThr, sizeof(float) * (ndiv + 1), cudaMemcpyHostToDevice, st));
    CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_stats.p, 0, kStatBytes, st));
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    if (phase == 1) CAB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));  // phase 2 may start once the prologue is in place
  }
  RsdArgs a{};
  a.g = grid_view(ctx);
  packet_range(ctx, &a.p0, &a.p1);
  a.range = slab_packet_range(ctx, false);
  a.slab = slab_info_device(ctx);
  if (phase == 1) {
    a.range = a.slab->r_int;  // device addresses of the SlabInfo's ranges
  } else if (phase == 2) {
    a.range = a.slab->r_blo;
    a.range_b = a.slab->r_bhi;
  }
  comm_push_targets(ctx, &a.push);
  const bool scatter = ctx->step_input_order && a.push.world == 0 && !ctx->slab;
  if (scatter) {  // a group of one: every result goes to its input index in this context's own arrays
    if (int rc = reserve(ctx, ctx->b_in_nrm, (size_t)std::max(n, 1) * sizeof(float4))) return rc;
    if (int rc = reserve(ctx, ctx->b_in_rsd, (size_t)std::max(n, 1) * sizeof(float2))) return rc;
    a.push.world = 1;
    a.push.layout = CAB_COMM_LAYOUT_INPUT_RANGES;
    a.push.n = std::max(n, 1);
    a.push.lo[0] = 0;
    a.push.lo[1] = n;
    a.push.nrm[0] = (float4*)ctx->b_in_nrm.p;
    a.push.rsd[0] = (float2*)ctx->b_in_rsd.p;
  }
  a.r = rf;
  a.r2 = r2;
  a.nrm = (const float4*)ctx->b_nrm.p;
  a.out = (float2*)ctx->b_rsd.p;
  a.rdif = (float*)ctx->b_rdif.p;
  a.rmin_in = ctx->slab ? nullptr : ctx->fuse_rmin_in;
  a.rmax_in = ctx->slab ? nullptr : ctx->fuse_rmax_in;
  a.thr_d2 = use_thr ? (const float*)ctx->b_thr_d2.p : nullptr;
  a.thr_idx = use_thr ? (const int*)ctx->b_thr_idx.p : nullptr;
  a.bin_thr = (const float*)ctx->b_stats2.p;
  a.ndiv = ndiv;
  a.flags = flags;
  a.bin_scale = (float)(ndiv / r);
  a.magic = 8388608.f;
  a.bin_scale_lo = (float)(ndiv / r * (1.0 - 1.0 / 262144.0));
  a.bin_scale_hi = (float)(ndiv / r * (1.0 + 1.0 / 262144.0));
  a.radius = r;
  a.plane_radius = plane_radius;

  a.stats = (unsigned long long*)ctx->b_stats.p;
  a.work_slot = slot;
  const size_t smem = (size_t)kWarpsPerBlock * sizeof(ChunkTile) + (size_t)kWarpsPerBlock * ndiv * kWarp * sizeof(float2) +
                      (size_t)((ndiv + 4) & ~3) * sizeof(float) + (size_t)kWarpsPerBlock * kWarp * sizeof(int);
  const int np = a.range ? std::max(1, (int)std::min<int64_t>(ctx->n_sorted, INT_MAX)) : a.p1 - a.p0;
  if (np > 0 && ctx->n_sorted > 0) {
    const unsigned blocks = (np + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int rc;
    if (!ctx->cfg.exact && (!use_thr || trunc_fast) && !legacy) {
      const size_t fsmem = (size_t)kWarpsPerBlock * sizeof(FastTile) + (size_t)kWarpsPerBlock * (ndiv + 1) * kRowBytes +
                           (size_t)kFastThr * sizeof(float) + 256 +
                           (trunc_fast ? (size_t)kWarpsPerBlock * kTruncCap * kWarp * sizeof(unsigned short) : 0);
      // neighbour counts kept by the last normals pass are this pass's counts if radius and rule were the same
      const bool counted = !trunc_fast && ctx->kcount_r == rf && ctx->kcount_valid;
      a.kcount = counted ? (const int*)ctx->b_kcount.p : nullptr;
      if (trunc_fast) {
        a.trunc_code = (const int*)ctx->b_thr_idx.p;
        a.trunc_scale = (float)kTruncBins / r2;
        a.skip = (const unsigned char*)ctx->b_thr_flag.p;
      }
      // unrolled over a whole chunk (8 groups of four candidates); the truncated variant, whose body is longer, half as far
      // (measured both ways: 10.1 against 10.6 ms, and 16.3 against 16.7 ms with max_nn = 150, on the 20 M-point room)
      auto kernel = trunc_fast ? rsd_fast_kernel<true, true, 4> : counted ? rsd_fast_kernel<false, false, 8> : rsd_fast_kernel<true, false, 8>;
      CAB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
      int per_sm = 1;
      CAB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerBlock * kWarp, fsmem));
      if (const char* g3 = std::getenv("CAB_FAST_PER_SM")) per_sm = std::min(per_sm, std::atoi(g3));
      const unsigned grid = std::min<unsigned>(blocks, (unsigned)std::max(per_sm, 1) * ctx->sm_count);
      kernel<<<grid, kWarpsPerBlock * kWarp, fsmem, ks>>>(a);
      CAB_LAUNCH_CHECK(ctx);
      rc = CAB_OK;
      if (trunc_fast) {  // the flagged packets: exact thresholds (computed by run_nn_hist for them) and the hit-compacting kernel
        CAB_CUDA(ctx, cudaMemsetAsync((unsigned long long*)ctx->b_stats.p + 2 * kStatSlots + slot, 0, 8, ks));  // packet work counter
        a.skip = nullptr;
        a.only = (const unsigned char*)ctx->b_thr_flag.p;
        rc = launch_rsd<false, true>(ctx, a, blocks, smem, ks);
      }
    } else if (ctx->cfg.exact) rc = use_thr ? launch_rsd<true, true>(ctx, a, blocks, smem, ks) : launch_rsd<true, false>(ctx, a, blocks, smem, ks);
    else rc = use_thr ? launch_rsd<false, true>(ctx, a, blocks, smem, ks) : launch_rsd<false, false>(ctx, a, blocks, smem, ks);
    if (rc) return rc;
  }
  if (phase == 1) return CAB_OK;  // the boundary packets and the epilogue follow in phase 2
  if (phase == 2) {               // join: the epilogue waits for both launches
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev_join, ks));
    CAB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
  }
  if (!ctx->slab && n > ctx->n_valid) {
    fill_invalid_rsd<<<(n - ctx->n_valid + 255) / 256, 256, 0, st>>>((float2*)ctx->b_rsd.p, (float*)ctx->b_rdif.p, ctx->n_valid, n,
                                                                    (float)plane_radius, (const int*)ctx->b_perm.p, a.rmin_in, a.rmax_in);
    CAB_LAUNCH_CHECK(ctx);
    if (scatter) {
      fill_invalid_input_order<<<(n - ctx->n_valid + 255) / 256, 256, 0, st>>>((const int*)ctx->b_perm.p, ctx->n_valid, n, (float)plane_radius,
                                                                              (float4*)ctx->b_in_nrm.p, (float2*)ctx->b_in_rsd.p);
      CAB_LAUNCH_CHECK(ctx);
    }
  }
  if (scatter) ctx->have_input_order = true;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->h_step + kStepStats1, ctx->b_stats.p, kStatBytes, cudaMemcpyDeviceToHost, st));
  ctx->have_rsd = true;
  if (ctx->defer_sync) return CAB_OK;
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return finish_pass_stats(ctx, 1);
}

}  // namespace cab
