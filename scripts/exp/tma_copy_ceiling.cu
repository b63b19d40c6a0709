// Experiment (not product code): what does the TMA load->smem->store pipeline of the headline kernel sustain
// when the consumers do no arithmetic?  Compares three staging geometries at N = 1e6 trajectories.
//   mode 0: three 2-D boxes per tile, 64 B / 64 B / 32 B rows, hardware swizzle (what the product kernel uses)
//   mode 1: two 1-D bulk copies per tile (8 KB + 12 KB), linear layout
//   mode 2: two 2-D boxes per tile over flat 128-B-row views, 128B swizzle
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../eigentrajectory_b200/csrc/et_common.cuh"
#include "../../eigentrajectory_b200/csrc/et_tma.cuh"
using namespace et;

constexpr int TILE = 128, STAGE = TILE * 160, THREADS = 160;
struct Maps { CUtensorMap a, b, c, oa, ob, oc; };

template <int NS, int MODE>
__global__ void __launch_bounds__(THREADS) copy_kernel(const __grid_constant__ Maps maps, const float* obs, const float* pred,
                                                       float* ro, float* rp, int n_tiles, int touch) {
  extern __shared__ uint8_t raw[];
  uint8_t* stages = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(stages + (size_t)NS * STAGE);
  uint64_t* done = full + NS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], 4); }
    fence_barrier_init();
  }
  __syncthreads();
  const int my = ((int)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (warp == 4) {
    if (lane == 0 && my > 0) {
      auto load = [&](int it) {
        const int s = it % NS; const int tile = blockIdx.x + it * gridDim.x; const int row0 = tile * TILE;
        uint8_t* st = stages + (size_t)s * STAGE;
        mbar_arrive_expect_tx(&full[s], STAGE);
        if (MODE == 0) {
          tma_load_2d(st, &maps.a, 0, row0, &full[s]);
          tma_load_2d(st + 8192, &maps.b, 0, row0, &full[s]);
          tma_load_2d(st + 16384, &maps.c, 16, row0, &full[s]);
        } else if (MODE == 1) {
          bulk_load(st, obs + (size_t)row0 * 16, 8192, &full[s]);
          bulk_load(st + 8192, pred + (size_t)row0 * 24, 12288, &full[s]);
        } else {
          tma_load_2d(st, &maps.a, 0, tile * 64, &full[s]);
          tma_load_2d(st + 8192, &maps.b, 0, tile * 96, &full[s]);
        }
      };
      auto store = [&](int it) {
        const int s = it % NS; const int tile = blockIdx.x + it * gridDim.x; const int row0 = tile * TILE;
        uint8_t* st = stages + (size_t)s * STAGE;
        if (MODE == 0) {
          tma_store_2d(&maps.oa, 0, row0, st);
          tma_store_2d(&maps.ob, 0, row0, st + 8192);
          tma_store_2d(&maps.oc, 16, row0, st + 16384);
        } else if (MODE == 1) {
          bulk_store(ro + (size_t)row0 * 16, st, 8192);
          bulk_store(rp + (size_t)row0 * 24, st + 8192, 12288);
        } else {
          tma_store_2d(&maps.oa, 0, tile * 64, st);
          tma_store_2d(&maps.ob, 0, tile * 96, st + 8192);
        }
        bulk_commit();
      };
      const int pre = my < NS ? my : NS;
      for (int it = 0; it < pre; ++it) load(it);
      for (int it = 0; it < my; ++it) {
        mbar_wait(&done[it % NS], (it / NS) & 1);
        store(it);
        if (it >= 1 && it - 1 + NS < my) { bulk_wait_read<1>(); load(it - 1 + NS); }
      }
      bulk_wait_all<0>();
    }
    return;
  }
  for (int it = 0; it < my; ++it) {
    const int s = it % NS;
    mbar_wait(&full[s], (it / NS) & 1);
    if (touch) {   // read-modify-write every float4 of my row-equivalent share (conflict pattern irrelevant here)
      float4* p = reinterpret_cast<float4*>(stages + (size_t)s * STAGE) + (warp * 32 + lane);
      for (int q = 0; q < 10; ++q) { float4 v = p[q * 128]; v.x += 1.f; p[q * 128] = v; }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);
  }
}

template <int NS, int MODE>
float run(int per_sm, const Maps& maps, std::vector<float*>& bufs, int n, int reps) {
  const size_t smem = 1024 + (size_t)NS * STAGE + 2 * NS * 8;
  cudaFuncSetAttribute(copy_kernel<NS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n_tiles = n / TILE;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, sum = 0;
  for (int r = 0; r < reps + 3; ++r) {
    cudaEventRecord(e0);
    copy_kernel<NS, MODE><<<sms * per_sm, THREADS, smem>>>(maps, bufs[0], bufs[1], bufs[2], bufs[3], n_tiles, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 3) { best = ms < best ? ms : best; sum += ms; }
  }
  cudaError_t e = cudaGetLastError();
  printf("mode %d NS %d blocks/SM %d smem %zu: avg %.2f us  best %.2f us  -> %.0f GB/s (320 MB)  [%s]\n", MODE, NS, per_sm, smem,
         1e3 * sum / reps, 1e3 * best, 320e6 / (sum / reps * 1e-3) / 1e9, cudaGetErrorString(e));
  return sum / reps;
}

int main(int argc, char** argv) {
  for (int n : {250112, 1000064, 4000000, 16000000}) {
    std::vector<float*> bufs(4);
    cudaMalloc(&bufs[0], (size_t)n * 64); cudaMalloc(&bufs[1], (size_t)n * 96);
    cudaMalloc(&bufs[2], (size_t)n * 64); cudaMalloc(&bufs[3], (size_t)n * 96);
    cudaMemset(bufs[0], 0, (size_t)n * 64); cudaMemset(bufs[1], 0, (size_t)n * 96);
    Maps m0;
    make_tensor_map_2d(&m0.a, bufs[0], n, 16, 64, 16, TILE, 64);
    make_tensor_map_2d(&m0.b, bufs[1], n, 24, 96, 16, TILE, 64);
    make_tensor_map_2d(&m0.c, bufs[1], n, 24, 96, 8, TILE, 32);
    make_tensor_map_2d(&m0.oa, bufs[2], n, 16, 64, 16, TILE, 64);
    make_tensor_map_2d(&m0.ob, bufs[3], n, 24, 96, 16, TILE, 64);
    make_tensor_map_2d(&m0.oc, bufs[3], n, 24, 96, 8, TILE, 32);
    printf("n = %d (%.0f MB moved)\n", n, n * 320e-6);
    float a = run<4, 0>(2, m0, bufs, n, 20);
    float b = run<4, 1>(2, m0, bufs, n, 20);
    printf("   mode0 %.0f GB/s, mode1 %.0f GB/s\n", n * 320.0 / (a * 1e-3) / 1e9, n * 320.0 / (b * 1e-3) / 1e9);
    for (auto p : bufs) cudaFree(p);
  }
  return 0;
}
