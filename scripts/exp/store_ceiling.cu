// Experiment (not product code): write bandwidth of the reconstruction kernel's store pattern with the arithmetic removed.
// out is (S=20, N, 24) floats; a unit of work is a tile of 32 pedestrians x 20 samples = 20 stores of 3 KB.
//   mode 0: per-warp 3 KB bulk stores (UBLKCP), 2 in flight per warp, block of 4 warps shares a tile (as reconstruct_fast)
//   mode 1: same but 4 slabs in flight per warp
//   mode 2: coalesced STG.128 from registers (each warp writes its 3 KB slab as 6 x 512 B)
//   mode 3: tile = 128 pedestrians per block, one 12 KB bulk store per sample issued by one thread after a block barrier
#include <cstdio>
#include <vector>
#include "../../eigentrajectory_b200/csrc/et_common.cuh"
#include "../../eigentrajectory_b200/csrc/et_tma.cuh"
using namespace et;
constexpr int S = 20, SLAB = 32 * 24;

template <int MODE, int NSLAB>
__global__ void __launch_bounds__(128) store_kernel(float* out, int64_t n, int64_t n_tiles) {
  extern __shared__ __align__(128) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slab = sm + warp * NSLAB * SLAB;
  for (int e = lane; e < NSLAB * SLAB; e += 32) slab[e] = (float)e;
  __syncthreads();
  if (MODE == 3) {
    float* big = sm;   // 2 x 12 KB
    int sel = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles / 4; tile += gridDim.x) {
      const int64_t n0 = tile * 128;
      for (int s = 0; s < S; ++s) {
        if (threadIdx.x == 0) bulk_wait_read<1>();
        __syncthreads();
        reinterpret_cast<float4*>(big + sel * 4 * SLAB)[threadIdx.x] = make_float4(1.f, 2.f, 3.f, (float)s);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_store(out + ((int64_t)s * n + n0) * 24, big + sel * 4 * SLAB, 4 * SLAB * 4); bulk_commit(); }
        sel ^= 1;
      }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();
    return;
  }
  int sel = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * 32;
    for (int s = warp; s < S; s += 4) {
      float* dst = out + ((int64_t)s * n + n0) * 24;
      if (MODE == 2) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int q = 0; q < 6; ++q) d4[q * 32 + lane] = make_float4(1.f, 2.f, 3.f, (float)s);
      } else {
        if (lane == 0) bulk_wait_read<NSLAB - 1>();
        __syncwarp();
        reinterpret_cast<float4*>(slab + sel * SLAB)[lane] = make_float4(1.f, 2.f, 3.f, (float)s);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { bulk_store(dst, slab + sel * SLAB, SLAB * 4); bulk_commit(); }
        sel = (sel + 1) % NSLAB;
      }
    }
  }
  if (MODE != 2 && lane == 0) bulk_wait_all<0>();
}

template <int MODE, int NSLAB>
void run(float* out, int64_t n, int per_sm, const char* what) {
  const int64_t n_tiles = n / 32;
  const size_t smem = (MODE == 3) ? 2 * 4 * SLAB * 4 : (size_t)4 * NSLAB * SLAB * 4;
  cudaFuncSetAttribute(store_kernel<MODE, NSLAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float sum = 0; const int reps = 20;
  for (int r = 0; r < reps + 3; ++r) {
    cudaEventRecord(e0);
    store_kernel<MODE, NSLAB><<<per_sm > 0 ? sms * per_sm : (int)n_tiles, 128, smem>>>(out, n, n_tiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 3) sum += ms;
  }
  const double bytes = (double)n * S * 96;
  printf("%-44s blocks/SM %d: %.1f us -> %.0f GB/s  [%s]\n", what, per_sm, 1e3 * sum / reps, bytes / (sum / reps * 1e-3) / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int64_t n = 200192;   // multiple of 128
  float* out; cudaMalloc(&out, (size_t)n * S * 96);
  printf("pure store of %.0f MB in the (S,N,24) pattern\n", n * S * 96e-6);
  run<0, 2>(out, n, 5, "per-warp 3 KB bulk stores, 2 slabs");
  run<0, 2>(out, n, 4, "per-warp 3 KB bulk stores, 2 slabs");
  run<1, 4>(out, n, 4, "per-warp 3 KB bulk stores, 4 slabs");
  run<2, 1>(out, n, 8, "coalesced STG.128 from registers");
  run<2, 1>(out, n, 0, "coalesced STG.128, one tile per block");
  run<3, 1>(out, n, 4, "per-block 12 KB bulk stores (tile 128)");
  run<3, 1>(out, n, 8, "per-block 12 KB bulk stores (tile 128)");
  return 0;
}
