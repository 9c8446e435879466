// Which arithmetic form of the normals inner loop is fastest on the B200?  Synthetic loop with the real
// instruction mix (3 broadcast LDS.128 per 4 candidates, distance test, predicated accumulation) in several
// variants of the distance computation.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o normals_mix normals_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 sq2(f32x2 a) { return fma2(a, a, 0ull); }

__device__ __forceinline__ void accum_pred(float d2, float r2, float dx, float dy, float dz, float& s1x, float& s1y, float& s1z,
                                           float& sxx, float& sxy, float& sxz, float& syy, float& syz, float& szz, int& k) {
  asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %10, %11;\n\t@p add.f32 %0, %0, %12;\n\t@p add.f32 %1, %1, %13;\n\t@p add.f32 %2, %2, %14;\n\t"
      "@p fma.rn.f32 %3, %12, %12, %3;\n\t@p fma.rn.f32 %4, %12, %13, %4;\n\t@p fma.rn.f32 %5, %12, %14, %5;\n\t"
      "@p fma.rn.f32 %6, %13, %13, %6;\n\t@p fma.rn.f32 %7, %13, %14, %7;\n\t@p fma.rn.f32 %8, %14, %14, %8;\n\t@p add.s32 %9, %9, 1;\n\t}"
      : "+f"(s1x), "+f"(s1y), "+f"(s1z), "+f"(sxx), "+f"(sxy), "+f"(sxz), "+f"(syy), "+f"(syz), "+f"(szz), "+r"(k)
      : "f"(d2), "f"(r2), "f"(dx), "f"(dy), "f"(dz));
}

// accumulate variants (second template parameter):
//  0: predicated 3 FADD + 6 FFMA + IADD (the kernel in the tree)
//  1: deltas zeroed for a miss (3 FSEL), then unpredicated 3 FADD + 6 FFMA, predicated IADD
//  2: as 1 with the three additions written as FFMA(d, 1, s)
//  3: predicated, the three additions written as FFMA(d, 1, s)
template <int ACC>
__device__ __forceinline__ void accum_var(float d2, float r2, float dx, float dy, float dz, float& s1x, float& s1y, float& s1z,
                                          float& sxx, float& sxy, float& sxz, float& syy, float& syz, float& szz, int& k) {
  if (ACC == 0) {
    accum_pred(d2, r2, dx, dy, dz, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, k);
  } else if (ACC == 1 || ACC == 2) {
    float ex, ey, ez;
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %4, %5;\n\tselp.f32 %0, %6, 0f00000000, p;\n\tselp.f32 %1, %7, 0f00000000, p;\n\t"
        "selp.f32 %2, %8, 0f00000000, p;\n\t@p add.s32 %3, %3, 1;\n\t}"
        : "=f"(ex), "=f"(ey), "=f"(ez), "+r"(k) : "f"(d2), "f"(r2), "f"(dx), "f"(dy), "f"(dz));
    if (ACC == 1) { s1x = __fadd_rn(s1x, ex); s1y = __fadd_rn(s1y, ey); s1z = __fadd_rn(s1z, ez); }
    else { s1x = __fmaf_rn(ex, 1.0f, s1x); s1y = __fmaf_rn(ey, 1.0f, s1y); s1z = __fmaf_rn(ez, 1.0f, s1z); }
    sxx = __fmaf_rn(ex, ex, sxx); sxy = __fmaf_rn(ex, ey, sxy); sxz = __fmaf_rn(ex, ez, sxz);
    syy = __fmaf_rn(ey, ey, syy); syz = __fmaf_rn(ey, ez, syz); szz = __fmaf_rn(ez, ez, szz);
  } else {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %10, %11;\n\t@p fma.rn.f32 %0, %12, %15, %0;\n\t@p fma.rn.f32 %1, %13, %15, %1;\n\t@p fma.rn.f32 %2, %14, %15, %2;\n\t"
        "@p fma.rn.f32 %3, %12, %12, %3;\n\t@p fma.rn.f32 %4, %12, %13, %4;\n\t@p fma.rn.f32 %5, %12, %14, %5;\n\t"
        "@p fma.rn.f32 %6, %13, %13, %6;\n\t@p fma.rn.f32 %7, %13, %14, %7;\n\t@p fma.rn.f32 %8, %14, %14, %8;\n\t@p add.s32 %9, %9, 1;\n\t}"
        : "+f"(s1x), "+f"(s1y), "+f"(s1z), "+f"(sxx), "+f"(sxy), "+f"(sxz), "+f"(syy), "+f"(syz), "+f"(szz), "+r"(k)
        : "f"(d2), "f"(r2), "f"(dx), "f"(dy), "f"(dz), "f"(r2 * 0.f + 1.0f));
  }
}

#define ITER 2048
// MODE 0: packed, no contraction (16 packed / 4 candidates)      -- the kernel in the tree
// MODE 1: scalar, no contraction (32 scalar)
// MODE 2: candidates 0,1 packed, 2,3 scalar
// MODE 3: packed with FMA chain d2 = fma(dz,dz,fma(dy,dy,dx*dx)) (12 packed)
// MODE 4: scalar with FMA chain (24 scalar)
template <int MODE, int ACC = 0>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  __shared__ __align__(16) float tx[32], ty[32], tz[32];
  if (threadIdx.x < 32) { tx[threadIdx.x] = seed * threadIdx.x; ty[threadIdx.x] = seed * 2 * threadIdx.x; tz[threadIdx.x] = seed * 3 * threadIdx.x; }
  __syncthreads();
  const float qx = seed * (threadIdx.x & 31), qy = qx * 2, qz = qx * 3, r2 = seed * seed * 40.f;
  const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
  float s1x = 0, s1y = 0, s1z = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
  int kk = 0;
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll 2
    for (int g4 = 0; g4 < 8; ++g4) {
      const float4 X = reinterpret_cast<const float4*>(tx)[g4], Y = reinterpret_cast<const float4*>(ty)[g4], Z = reinterpret_cast<const float4*>(tz)[g4];
      float d2a, d2b, xa, xb, ya, yb, za, zb;
#define PACKED(X0, X1, Y0, Y1, Z0, Z1, FMA)                                                                    \
  {                                                                                                            \
    const f32x2 dx = sub2(pack2(X0, X1), qx2), dy = sub2(pack2(Y0, Y1), qy2), dz = sub2(pack2(Z0, Z1), qz2);   \
    const f32x2 d2 = FMA ? fma2(dz, dz, fma2(dy, dy, sq2(dx))) : add2(add2(sq2(dx), sq2(dy)), sq2(dz));        \
    unpack2(d2, d2a, d2b); unpack2(dx, xa, xb); unpack2(dy, ya, yb); unpack2(dz, za, zb);                      \
  }
#define SCALAR(X0, Y0, Z0, FMA, D2, DX, DY, DZ)                                                                \
  {                                                                                                            \
    DX = __fsub_rn(X0, qx); DY = __fsub_rn(Y0, qy); DZ = __fsub_rn(Z0, qz);                                    \
    D2 = FMA ? __fmaf_rn(DZ, DZ, __fmaf_rn(DY, DY, __fmul_rn(DX, DX)))                                         \
             : __fadd_rn(__fadd_rn(__fmul_rn(DX, DX), __fmul_rn(DY, DY)), __fmul_rn(DZ, DZ));                  \
  }
#define ACC2 accum_var<ACC>(d2a, r2, xa, ya, za, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, kk); \
             accum_var<ACC>(d2b, r2, xb, yb, zb, s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz, kk);
      if (MODE >= 5) {
        // mixed forms: which part of the distance test runs packed
        //  5: subtractions packed (6 per 4 candidates), squares and sums scalar (20)
        //  6: subtractions and squares packed (12), sums scalar (8)
        //  7: subtractions scalar (12), squares and sums packed (10)
#define MIXED(X0, X1, Y0, Y1, Z0, Z1)                                                                          \
  {                                                                                                            \
    if (MODE == 5) {                                                                                           \
      const f32x2 dx = sub2(pack2(X0, X1), qx2), dy = sub2(pack2(Y0, Y1), qy2), dz = sub2(pack2(Z0, Z1), qz2); \
      unpack2(dx, xa, xb); unpack2(dy, ya, yb); unpack2(dz, za, zb);                                           \
      d2a = __fadd_rn(__fadd_rn(__fmul_rn(xa, xa), __fmul_rn(ya, ya)), __fmul_rn(za, za));                     \
      d2b = __fadd_rn(__fadd_rn(__fmul_rn(xb, xb), __fmul_rn(yb, yb)), __fmul_rn(zb, zb));                     \
    } else if (MODE == 6) {                                                                                    \
      const f32x2 dx = sub2(pack2(X0, X1), qx2), dy = sub2(pack2(Y0, Y1), qy2), dz = sub2(pack2(Z0, Z1), qz2); \
      float p0, p1, q0, q1, r0, r1;                                                                            \
      unpack2(sq2(dx), p0, p1); unpack2(sq2(dy), q0, q1); unpack2(sq2(dz), r0, r1);                            \
      unpack2(dx, xa, xb); unpack2(dy, ya, yb); unpack2(dz, za, zb);                                           \
      d2a = __fadd_rn(__fadd_rn(p0, q0), r0); d2b = __fadd_rn(__fadd_rn(p1, q1), r1);                          \
    } else {                                                                                                   \
      xa = __fsub_rn(X0, qx); xb = __fsub_rn(X1, qx); ya = __fsub_rn(Y0, qy); yb = __fsub_rn(Y1, qy);          \
      za = __fsub_rn(Z0, qz); zb = __fsub_rn(Z1, qz);                                                          \
      const f32x2 dx = pack2(xa, xb), dy = pack2(ya, yb), dz = pack2(za, zb);                                  \
      unpack2(add2(add2(sq2(dx), sq2(dy)), sq2(dz)), d2a, d2b);                                                \
    }                                                                                                          \
  }
        MIXED(X.x, X.y, Y.x, Y.y, Z.x, Z.y) ACC2
        MIXED(X.z, X.w, Y.z, Y.w, Z.z, Z.w) ACC2
      } else if (MODE == 0 || MODE == 3) {
        PACKED(X.x, X.y, Y.x, Y.y, Z.x, Z.y, MODE == 3) ACC2
        PACKED(X.z, X.w, Y.z, Y.w, Z.z, Z.w, MODE == 3) ACC2
      } else if (MODE == 1 || MODE == 4) {
        SCALAR(X.x, Y.x, Z.x, MODE == 4, d2a, xa, ya, za) SCALAR(X.y, Y.y, Z.y, MODE == 4, d2b, xb, yb, zb) ACC2
        SCALAR(X.z, Y.z, Z.z, MODE == 4, d2a, xa, ya, za) SCALAR(X.w, Y.w, Z.w, MODE == 4, d2b, xb, yb, zb) ACC2
      } else {
        PACKED(X.x, X.y, Y.x, Y.y, Z.x, Z.y, false) ACC2
        SCALAR(X.z, Y.z, Z.z, false, d2a, xa, ya, za) SCALAR(X.w, Y.w, Z.w, false, d2b, xb, yb, zb) ACC2
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s1x + s1y + s1z + sxx + sxy + sxz + syy + syz + szz + kk;
}

template <int MODE, int ACC = 0>
void run(const char* name, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm;
  float* out;
  cudaMalloc(&out, blocks * 256 * 4);
  k<MODE, ACC><<<blocks, 256>>>(out, 0.01f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE, ACC><<<blocks, 256>>>(out, 0.01f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  // groups of 4 candidates per SMSP: blocks_per_sm * 8 warps / 4 SMSPs * ITER * 8
  const double groups = (double)blocks_per_sm * 2 * ITER * 8;
  printf("%-44s %d blocks/SM: %7.3f ms -> %6.1f cycles per 4 candidates per SMSP-warp slot\n", name, blocks_per_sm, best,
         best * 1.965e6 / groups);
  cudaFree(out);
}

int main() {
  run<0, 3>("packed dist (16P), all-FFMA acc", 4);
  run<1, 3>("scalar dist (32S), all-FFMA acc", 4);
  run<2, 3>("half packed / half scalar, all-FFMA acc", 4);
  run<5, 3>("sub packed (6P) + 20S, all-FFMA acc", 4);
  run<6, 3>("sub+sq packed (12P) + 8S, all-FFMA acc", 4);
  run<7, 3>("sub scalar (12S) + sq/add packed (10P)", 4);
  run<0, 0>("packed dist, predicated acc (old tree)", 4);
  return 0;
}
