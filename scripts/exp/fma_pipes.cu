// Micro-benchmark: issue rate of packed (FFMA2) and scalar (FFMA) fp32 multiply-adds and of mixes with ALU work on
// B200 -- decides how the k-means similarity scan should split its multiply-adds.   nvcc -arch=sm_100a -O3 fma_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int N2, int N1, int NALU>      // per inner step: N2 packed FMAs, N1 scalar FMAs, NALU compare+select pairs
__global__ void __launch_bounds__(384, 1) kern(float* out, float x, int iters, long long* cycles) {
  u64 acc2[8]; float acc1[8]; float best = -1e30f; int lab = 0;
  for (int i = 0; i < 8; ++i) { acc2[i] = (u64)threadIdx.x * 3 + i; acc1[i] = threadIdx.x + i; }
  u64 b2 = ((u64)__float_as_uint(x) << 32) | __float_as_uint(x * 0.5f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < N2; ++i) acc2[(r + i) & 7] = fma2(acc2[(r + i) & 7], b2, acc2[(r + i + 1) & 7]);
#pragma unroll
      for (int i = 0; i < N1; ++i) acc1[(r + i) & 7] = fma1(acc1[(r + i) & 7], x, acc1[(r + i + 3) & 7]);
#pragma unroll
      for (int i = 0; i < NALU; ++i) { float y = acc1[(r + i) & 7]; if (y > best) { best = y; lab = r * 8 + i + it; } }
    }
  }
  long long t1 = clock64();
  float s = best + lab;
  for (int i = 0; i < 8; ++i) s += acc1[i] + __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int N2, int N1, int NALU>
void run(const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  kern<N2, N1, NALU><<<148, 384>>>(out, 1.0001f, iters, cyc);
  kern<N2, N1, NALU><<<148, 384>>>(out, 1.0001f, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double steps = (double)iters * 8;
  const double per_warp_step = (double)c / steps;                 // cycles per inner step, 3 warps per scheduler
  printf("%-34s %7.2f cycles per step per scheduler-warp-triple -> %.2f cyc per FFMA2, %.2f per FFMA, fp32 FMA lanes/clk/SM %.1f\n", name,
         per_warp_step, N2 ? per_warp_step / (3.0 * N2) : 0.0, N1 ? per_warp_step / (3.0 * N1) : 0.0,
         (N2 * 2 + N1) * 32.0 * 12 / per_warp_step);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<8, 0, 0>("FFMA2 only");
  run<0, 8, 0>("FFMA only");
  run<4, 8, 0>("FFMA2 : FFMA = 1 : 2");
  run<4, 4, 0>("FFMA2 : FFMA = 1 : 1");
  run<6, 2, 0>("FFMA2 : FFMA = 3 : 1");
  run<8, 0, 4>("FFMA2 8 + 4 cmp/sel");
  run<0, 8, 4>("FFMA 8 + 4 cmp/sel");
  run<4, 4, 4>("FFMA2 4 + FFMA 4 + 4 cmp/sel");
  return 0;
}
