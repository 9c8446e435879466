// Micro-benchmark of the FP32 issue / pipe rates that matter for the traversal kernels (sm_100a):
// scalar FFMA vs packed FFMA2 / FADD2, FSETP, FSET, SHFL, LDS, shared atomics.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
#define ITER 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
  __shared__ float sh[1024];
  __shared__ unsigned shu[1024];
  float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  f32x2 p0, p1, p2, p3, p4, p5, p6, p7, m;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
  p4 = p0; p5 = p1; p6 = p2; p7 = p3;
  asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(seed * 0.5f), "f"(seed * 0.25f));
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sh[i] = i; shu[i] = i; }
  __syncthreads();
  const float c = seed * 0.999f;
  int idx = threadIdx.x & 31;
  unsigned pr = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {  // 8 independent scalar FFMA
      a0 = fmaf(a0, c, c); a1 = fmaf(a1, c, c); a2 = fmaf(a2, c, c); a3 = fmaf(a3, c, c);
      a4 = fmaf(a4, c, c); a5 = fmaf(a5, c, c); a6 = fmaf(a6, c, c); a7 = fmaf(a7, c, c);
    } else if (MODE == 1) {  // 8 independent FFMA2
#define F2(p) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(m));
      F2(p0) F2(p1) F2(p2) F2(p3) F2(p4) F2(p5) F2(p6) F2(p7)
    } else if (MODE == 2) {  // 8 independent FADD2
#define A2(p) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(m));
      A2(p0) A2(p1) A2(p2) A2(p3) A2(p4) A2(p5) A2(p6) A2(p7)
    } else if (MODE == 3) {  // 4 FFMA2 + 4 scalar FFMA
      F2(p0) F2(p1) F2(p2) F2(p3)
      a0 = fmaf(a0, c, c); a1 = fmaf(a1, c, c); a2 = fmaf(a2, c, c); a3 = fmaf(a3, c, c);
    } else if (MODE == 4) {  // 8 scalar FADD
      a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
    } else if (MODE == 5) {  // 8 FSETP + predicated int add (ALU)
#define SP(x) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.s32 %0, %0, 1;}" : "+r"(pr) : "f"(x), "f"(c));
      SP(a0) SP(a1) SP(a2) SP(a3) SP(a4) SP(a5) SP(a6) SP(a7)
      a0 += 1.f;
    } else if (MODE == 6) {  // 8 SHFL.IDX
      a0 = __shfl_sync(~0u, a0, idx); a1 = __shfl_sync(~0u, a1, idx); a2 = __shfl_sync(~0u, a2, idx); a3 = __shfl_sync(~0u, a3, idx);
      a4 = __shfl_sync(~0u, a4, idx); a5 = __shfl_sync(~0u, a5, idx); a6 = __shfl_sync(~0u, a6, idx); a7 = __shfl_sync(~0u, a7, idx);
    } else if (MODE == 7) {  // 8 LDS.32, lane-varying conflict-free
      a0 += sh[idx]; a1 += sh[idx + 32]; a2 += sh[idx + 64]; a3 += sh[idx + 96];
      a4 += sh[idx + 128]; a5 += sh[idx + 160]; a6 += sh[idx + 192]; a7 += sh[idx + 224];
      idx = (idx + 1) & 31;
    } else if (MODE == 8) {  // 8 shared atomicMin (no return), conflict-free
      unsigned v = __float_as_uint(a0);
      atomicMin(&shu[idx], v); atomicMin(&shu[idx + 32], v); atomicMin(&shu[idx + 64], v); atomicMin(&shu[idx + 96], v);
      atomicMax(&shu[idx + 128], v); atomicMax(&shu[idx + 160], v); atomicMax(&shu[idx + 192], v); atomicMax(&shu[idx + 224], v);
      a0 += 1.f;
    } else if (MODE == 9) {  // 4 x (LDS.64 + STS.64) lane-private
      float2* s2 = reinterpret_cast<float2*>(sh);
      float2 v0 = s2[idx], v1 = s2[idx + 32], v2 = s2[idx + 64], v3 = s2[idx + 96];
      v0.x += 1.f; v1.x += 1.f; v2.x += 1.f; v3.x += 1.f;
      s2[idx] = v0; s2[idx + 32] = v1; s2[idx + 64] = v2; s2[idx + 96] = v3;
    } else if (MODE == 10) {  // 8 LDS.128 broadcast (all lanes the same address)
      const float4* s4 = reinterpret_cast<const float4*>(sh);
      float4 v0 = s4[it & 15], v1 = s4[(it + 1) & 15], v2 = s4[(it + 2) & 15], v3 = s4[(it + 3) & 15];
      float4 v4 = s4[(it + 4) & 15], v5 = s4[(it + 5) & 15], v6 = s4[(it + 6) & 15], v7 = s4[(it + 7) & 15];
      a0 += v0.x + v0.w; a1 += v1.y; a2 += v2.z; a3 += v3.w; a4 += v4.x; a5 += v5.y; a6 += v6.z; a7 += v7.w;
    } else if (MODE == 11) {  // 8 FSET.BF (float compare -> 1.0/0.0)
#define FS(x) asm volatile("set.le.f32.f32 %0, %0, %1;" : "+f"(x) : "f"(c));
      FS(a0) FS(a1) FS(a2) FS(a3) FS(a4) FS(a5) FS(a6) FS(a7)
    } else if (MODE == 12) {  // 8 predicated scalar FFMA (predicate true on half the lanes)
#define PF(x) asm volatile("{.reg .pred q; setp.lt.u32 q, %2, 16; @q fma.rn.f32 %0, %0, %1, %1;}" : "+f"(x) : "f"(c), "r"(idx));
      PF(a0) PF(a1) PF(a2) PF(a3) PF(a4) PF(a5) PF(a6) PF(a7)
    }
  }
  long long t1 = clock64();
  float r0, r1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(p0));
  float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + r0 + r1 + pr + shu[threadIdx.x] + sh[threadIdx.x];
  f32x2 q = p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(q & 0xff);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * 256 * 4);
  cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, 256>>>(out, cyc, 1.0001f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, cyc, 1.0001f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[4];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // warp-instructions per SMSP: blocks_per_sm * 8 warps / 4 SMSP * ITER * ops
  const double winst = (double)blocks_per_sm * 2 * ITER * ops_per_iter;
  printf("%-28s blocks/SM %d: %8.3f ms, block cycles %lld -> %.2f cycles per warp-instruction per SMSP\n", name, blocks_per_sm, ms,
         h[0], (double)h[0] / winst);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int b : {1, 4}) {
    if (b == 1) {
      run<0>("FFMA x8", 8, 1); run<1>("FFMA2 x8", 8, 1); run<2>("FADD2 x8", 8, 1); run<3>("FFMA2 x4 + FFMA x4", 8, 1);
      run<4>("FADD x8", 8, 1); run<5>("FSETP+@IADD x8", 16, 1); run<6>("SHFL.IDX x8", 8, 1); run<7>("LDS.32 x8", 8, 1);
      run<8>("ATOMS min/max x8", 8, 1); run<9>("LDS.64+STS.64 x4", 8, 1); run<10>("LDS.128 bcast x8", 8, 1);
      run<11>("FSET.BF x8", 8, 1); run<12>("@p FFMA x8", 8, 1);
    } else {
      run<0>("FFMA x8", 8, 4); run<1>("FFMA2 x8", 8, 4); run<2>("FADD2 x8", 8, 4); run<3>("FFMA2 x4 + FFMA x4", 8, 4);
      run<4>("FADD x8", 8, 4); run<5>("FSETP+@IADD x8", 16, 4); run<6>("SHFL.IDX x8", 8, 4); run<7>("LDS.32 x8", 8, 4);
      run<8>("ATOMS min/max x8", 8, 4); run<9>("LDS.64+STS.64 x4", 8, 4); run<10>("LDS.128 bcast x8", 8, 4);
      run<11>("FSET.BF x8", 8, 4); run<12>("@p FFMA x8", 8, 4);
    }
  }
  return 0;
}
