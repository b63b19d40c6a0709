// Fused min-over-S ADE / FDE (reference: utils/metrics.py:73-102, both functions in one pass).
//
// pred (S,N,T,2) is read exactly once: 96 B per (sample, pedestrian) row, 2024 algorithmic bytes per
// pedestrian at S=20, T=12.  Each warp owns 32 pedestrians and streams the S slabs
// pred[s, n0:n0+32] (contiguous 32*T*8 bytes) through a warp-private ring of 1-D bulk copies; the
// running minima live in registers.  HBM-bound; no block-wide barrier after set-up.  One tile per warp, 16-warp
// blocks: the hardware block scheduler balances the load.
#include "et_common.cuh"
#include "et_tma.cuh"

namespace et {

__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int T, int NSTAGE, int WARPS>
struct AdeSmem {
  static constexpr int SLAB_FLOATS = 32 * 2 * T;
  static constexpr size_t bytes = 128 + (size_t)WARPS * NSTAGE * SLAB_FLOATS * 4 + (size_t)WARPS * NSTAGE * 8;
};

template <int T, int NSTAGE, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) ade_fde_fast(const float* __restrict__ pred, const float* __restrict__ gt,
                                                           int s_total, int64_t n, int64_t n_tiles,
                                                           float* __restrict__ ade, float* __restrict__ fde,
                                                           int32_t* __restrict__ argmin_fde) {
  using L = AdeSmem<T, NSTAGE, WARPS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(base) + (size_t)warp * NSTAGE * L::SLAB_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)WARPS * NSTAGE * L::SLAB_FLOATS * 4) + warp * NSTAGE;
  if (lane < NSTAGE) mbar_init(&bars[lane], 1);
  __syncwarp();
  if (lane == 0) fence_barrier_init();
  __syncwarp();

  const int64_t wstride = (int64_t)gridDim.x * WARPS;
  const int64_t first = (int64_t)blockIdx.x * WARPS + warp;
  if (first >= n_tiles) return;
  const int64_t row_bytes_all = n * 2 * T;          // floats between consecutive samples

  // producer cursor (lane 0): next slab to request = (p_tile, p_s) into ring slot p_slot; all incremental, no divisions
  int64_t p_tile = first;
  int p_s = 0, p_slot = 0;
  auto issue = [&]() {
    const int64_t n0 = p_tile * 32;
    const int rows = (int)((n - n0) < 32 ? (n - n0) : 32);
    const uint32_t bytes = (uint32_t)(rows * 2 * T * 4);
    mbar_arrive_expect_tx(&bars[p_slot], bytes);
    bulk_load(ring + p_slot * L::SLAB_FLOATS, pred + (int64_t)p_s * row_bytes_all + n0 * 2 * T, bytes, &bars[p_slot]);
    if (++p_s == s_total) { p_s = 0; p_tile += wstride; }
    p_slot = (p_slot + 1 == NSTAGE) ? 0 : p_slot + 1;
  };
  if (lane == 0) {
#pragma unroll 1
    for (int q = 0; q < NSTAGE && p_tile < n_tiles; ++q) issue();
  }

  int slot = 0;
  uint32_t parity = 0;
  for (int64_t tile = first; tile < n_tiles; tile += wstride) {
    const int64_t i = tile * 32 + lane;
    const bool live = i < n;
    float g[2 * T];
    {
      const float4* p = reinterpret_cast<const float4*>(gt + (live ? i : n - 1) * 2 * T);
#pragma unroll
      for (int q = 0; q < 2 * T / 4; ++q) {
        const float4 v = __ldg(p + q);
        g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
      }
    }
    float best_ade = 0.f, best_fde = 0.f;
    int best_idx = 0;
#pragma unroll 1
    for (int s = 0; s < s_total; ++s) {
      mbar_wait(&bars[slot], parity);
      const float4* sl = reinterpret_cast<const float4*>(ring + slot * L::SLAB_FLOATS + lane * 2 * T);
      float4 v[2 * T / 4];
#pragma unroll
      for (int q = 0; q < 2 * T / 4; ++q) v[q] = sl[q];
      __syncwarp();   // every lane holds its row in registers: the slot can be refilled right away
      if (lane == 0 && p_tile < n_tiles) issue();
      // displacement norms: sqrt.approx (<= 1 ulp); four partial sums for instruction-level parallelism
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      float last = 0.f;
#pragma unroll
      for (int q = 0; q < 2 * T / 4; ++q) {
        const float dx0 = v[q].x - g[4 * q], dy0 = v[q].y - g[4 * q + 1];
        const float dx1 = v[q].z - g[4 * q + 2], dy1 = v[q].w - g[4 * q + 3];
        const float d0 = fast_sqrt(fmaf(dy0, dy0, dx0 * dx0));
        const float d1 = fast_sqrt(fmaf(dy1, dy1, dx1 * dx1));
        part[(2 * q) & 3] += d0;
        part[(2 * q + 1) & 3] += d1;
        last = d1;
      }
      const float sum = (part[0] + part[1]) + (part[2] + part[3]);
      const float a = sum / (float)T;
      // torch.min semantics: first minimum wins, NaN propagates
      if (s == 0 || a < best_ade || (a != a && best_ade == best_ade)) best_ade = a;
      if (s == 0 || last < best_fde || (last != last && best_fde == best_fde)) { best_fde = last; best_idx = s; }
      if (++slot == NSTAGE) { slot = 0; parity ^= 1u; }
    }
    if (live) {
      ade[i] = best_ade;
      fde[i] = best_fde;
      if (argmin_fde) argmin_fde[i] = best_idx;
    }
  }
}

// Any T: one thread per pedestrian, direct global reads.
__global__ void ade_fde_generic(const float* __restrict__ pred, const float* __restrict__ gt, int s_total, int64_t n,
                                int t, float* __restrict__ ade, float* __restrict__ fde,
                                int32_t* __restrict__ argmin_fde) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2* g = reinterpret_cast<const float2*>(gt) + i * t;
  float best_ade = 0.f, best_fde = 0.f;
  int best_idx = 0;
  for (int s = 0; s < s_total; ++s) {
    const float2* p = reinterpret_cast<const float2*>(pred) + ((int64_t)s * n + i) * t;
    float sum = 0.f, last = 0.f;
    for (int q = 0; q < t; ++q) {
      const float2 a = __ldg(p + q), b = __ldg(g + q);
      const float dx = a.x - b.x, dy = a.y - b.y;
      last = fast_sqrt(fmaf(dy, dy, dx * dx));
      sum += last;
    }
    const float a = sum / (float)t;
    if (s == 0 || a < best_ade || (a != a && best_ade == best_ade)) best_ade = a;
    if (s == 0 || last < best_fde || (last != last && best_fde == best_fde)) { best_fde = last; best_idx = s; }
  }
  ade[i] = best_ade;
  fde[i] = best_fde;
  if (argmin_fde) argmin_fde[i] = best_idx;
}

template <int WARPS, int NSTAGE, int MINB = 1>
static int launch_ade(const float* pred, const float* gt, int s, int64_t n, int64_t n_tiles, float* ade, float* fde,
                      int32_t* argmin_fde, int blocks_per_sm, cudaStream_t st) {
  using L = AdeSmem<12, NSTAGE, WARPS>;
  auto kern = ade_fde_fast<12, NSTAGE, WARPS, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "ade_fde_fast: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int64_t grid = (n_tiles + WARPS - 1) / WARPS;
  // blocks_per_sm > 0: persistent grid of that many blocks per SM; 0: one tile per warp, the hardware block scheduler
  // balances the load (best when a tile is a sizeable unit of work, as here: S slabs of 3 KB)
  const int64_t cap = blocks_per_sm > 0 ? (int64_t)sm_count() * blocks_per_sm : ((int64_t)1 << 30);
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, WARPS * 32, L::bytes, st>>>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde);
  return check_launch("ade_fde_fast");
}

}  // namespace et

using namespace et;

extern "C" int et_ade_fde(const float* pred, const float* gt, int s, int64_t n, int t, float* ade, float* fde,
                          int32_t* argmin_fde, et_stream_t stream) {
  ET_REQUIRE(n >= 0 && s >= 1 && t >= 1, ET_ERR_BADARG, "et_ade_fde: bad shape (S=%d, N=%lld, T=%d)", s, (long long)n, t);
  ET_REQUIRE((pred && gt && ade && fde) || n == 0, ET_ERR_BADARG, "et_ade_fde: null pointer");
  ET_REQUIRE(aligned16(pred) && aligned16(gt), ET_ERR_ALIGN, "et_ade_fde: pred / gt must be 16-byte aligned");
  if (n == 0) return ET_OK;
  cudaStream_t st = as_stream(stream);
  if (t == 12 && n >= 32) {
    const int64_t n_tiles = (n + 31) / 32;
    // <warps per block, ring depth, min blocks/SM>(..., blocks per SM of a persistent grid | 0 = one tile per warp).
    // Measured at S = 20, N = 2e5 on B200: 79 us for the default vs 87-93 us for the persistent shapes.
    switch (tune_get(ET_TUNE_ADE_CONFIG)) {
      case 1: return launch_ade<8, 4, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, 2, st);
      case 2: return launch_ade<16, 2, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, 2, st);
      case 3: return launch_ade<8, 3, 3>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, 0, st);
      default: return launch_ade<16, 2, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, 0, st);
    }
  }
  ade_fde_generic<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pred, gt, s, n, t, ade, fde, argmin_fde);
  return check_launch("ade_fde_generic");
}
