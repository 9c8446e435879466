// Micro-benchmark of the fp64 pipe and the fp32 -> fp64 conversion on the B200 (sm_100a): what the exact-mode
// accumulation of the normals kernel costs per hit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64 fp64.cu ; run: ./fp64
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, long long* cyc, double seed) {
  __shared__ double shd[512];
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  float f0 = (float)a0, f1 = (float)a1, f2 = (float)a2, f3 = (float)a3, f4 = (float)a4, f5 = (float)a5, f6 = (float)a6, f7 = (float)a7;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) shd[i] = i;
  __syncthreads();
  const double c = seed * 0.999;
  const float cf = (float)c;
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (MODE == 0) {  // 8 DFMA
      a0 = fma(a0, c, c); a1 = fma(a1, c, c); a2 = fma(a2, c, c); a3 = fma(a3, c, c);
      a4 = fma(a4, c, c); a5 = fma(a5, c, c); a6 = fma(a6, c, c); a7 = fma(a7, c, c);
    } else if (MODE == 1) {  // 8 DADD
      a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
    } else if (MODE == 2) {  // 8 x (FADD + F2F.F64.F32 + DADD)
      f0 += cf; f1 += cf; f2 += cf; f3 += cf; f4 += cf; f5 += cf; f6 += cf; f7 += cf;
      a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3;
      a4 += (double)f4; a5 += (double)f5; a6 += (double)f6; a7 += (double)f7;
    } else if (MODE == 3) {  // 8 x (FADD + DADD of a double kept in a register): MODE 2 without the conversion
      f0 += cf; f1 += cf; f2 += cf; f3 += cf; f4 += cf; f5 += cf; f6 += cf; f7 += cf;
      a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
    } else if (MODE == 4) {  // 8 LDS.64, lane-varying, conflict-free + DADD
      a0 += shd[idx]; a1 += shd[idx + 32]; a2 += shd[idx + 64]; a3 += shd[idx + 96];
      a4 += shd[idx + 128]; a5 += shd[idx + 160]; a6 += shd[idx + 192]; a7 += shd[idx + 224];
      idx = (idx + 1) & 31;
    } else if (MODE == 5) {  // the exact-mode hit: 3 cvt + 3 DADD + 6 DFMA
      const double ex = (double)f0, ey = (double)f1, ez = (double)f2;
      f0 += cf; f1 += cf; f2 += cf;
      a0 += ex; a1 += ey; a2 += ez;
      a3 = fma(ex, ex, a3); a4 = fma(ex, ey, a4); a5 = fma(ex, ez, a5); a6 = fma(ey, ey, a6); a7 = fma(ey, ez, a7);
      a0 = fma(ez, ez, a0);
    } else if (MODE == 6) {  // the same hit from fp64 operands: 3 LDS.64 + 3 DADD (sub) + 3 DADD + 6 DFMA
      const double ex = shd[idx] - c, ey = shd[idx + 32] - c, ez = shd[idx + 64] - c;
      idx = (idx + 1) & 31;
      a0 += ex; a1 += ey; a2 += ez;
      a3 = fma(ex, ex, a3); a4 = fma(ex, ey, a4); a5 = fma(ex, ez, a5); a6 = fma(ey, ey, a6); a7 = fma(ey, ez, a7);
      a0 = fma(ez, ez, a0);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + idx;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * blocks_per_sm;
  double* out;
  long long* cyc;
  cudaMalloc(&out, blocks * 256 * 8);
  cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, 256>>>(out, cyc, 1.0001);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, cyc, 1.0001);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[1];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double groups = (double)blocks_per_sm * 2 * ITER;  // loop iterations per SMSP (8 warps per block over 4 SMSPs)
  printf("%-44s blocks/SM %d: %8.3f ms, %.1f cycles per loop iteration per SMSP (%d listed ops -> %.2f cycles each)\n", name,
         blocks_per_sm, ms, (double)h[0] / groups, ops_per_iter, (double)h[0] / groups / ops_per_iter);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int b : {1, 4}) {
    if (b == 1) {
      run<0>("DFMA x8", 8, 1); run<1>("DADD x8", 8, 1); run<2>("(FADD + F2F.F64.F32 + DADD) x8", 8, 1);
      run<3>("(FADD + DADD) x8", 8, 1); run<4>("(LDS.64 + DADD) x8", 8, 1); run<5>("hit: 3 cvt + 3 DADD + 6 DFMA", 1, 1);
      run<6>("hit: 3 LDS.64 + 6 DADD + 6 DFMA", 1, 1);
    } else {
      run<0>("DFMA x8", 8, 4); run<1>("DADD x8", 8, 4); run<2>("(FADD + F2F.F64.F32 + DADD) x8", 8, 4);
      run<3>("(FADD + DADD) x8", 8, 4); run<4>("(LDS.64 + DADD) x8", 8, 4); run<5>("hit: 3 cvt + 3 DADD + 6 DFMA", 1, 4);
      run<6>("hit: 3 LDS.64 + 6 DADD + 6 DFMA", 1, 4);
    }
  }
  return 0;
}
