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

// Temporal correlation coefficient of one pedestrian (utils/metrics.py:112-129): Pearson correlation over the T frames
// between the best-FDE sample and the ground truth, separately for x and y, clamped to [-1, 1], NaN -> 0, averaged.
// Same operation order as the reference: centred series, (factor * centred) . centred, divide by both std-devs.
template <int T>
__device__ __forceinline__ float tcc_of(const float (&p)[2 * T], const float (&g)[2 * T]) {
  const float factor = 1.0f / (float)(T - 1);
  float out = 0.f;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float mp = 0.f, mg = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) { mp += p[2 * t + c]; mg += g[2 * t + c]; }
    mp /= (float)T;
    mg /= (float)T;
    float cpg = 0.f, cpp = 0.f, cgg = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float dp = p[2 * t + c] - mp, dg = g[2 * t + c] - mg;
      cpg = fmaf(factor * dp, dg, cpg);
      cpp = fmaf(factor * dp, dp, cpp);
      cgg = fmaf(factor * dg, dg, cgg);
    }
    const float raw = cpg / sqrtf(cpp) / sqrtf(cgg);
    out += (raw == raw) ? fminf(fmaxf(raw, -1.f), 1.f) : 0.f;   // clamp, NaN (constant series) -> 0
  }
  return out / 2.0f;
}

template <int T, int NSTAGE, int WARPS>
struct AdeSmem {
  static constexpr int SLAB_FLOATS = 32 * 2 * T;
  static constexpr size_t bytes = 128 + (size_t)WARPS * NSTAGE * SLAB_FLOATS * 4 + (size_t)WARPS * NSTAGE * 8;
};

template <int T, int NSTAGE, int WARPS, int MINB, bool TCC>
__global__ void __launch_bounds__(WARPS * 32, MINB) ade_fde_fast(const float* __restrict__ pred, const float* __restrict__ gt,
                                                           int s_total, int64_t n, int64_t n_tiles,
                                                           float* __restrict__ ade, float* __restrict__ fde,
                                                           int32_t* __restrict__ argmin_fde, float* __restrict__ tcc) {
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
    float4 best_row[TCC ? 2 * T / 4 : 1];     // the best-FDE sample's trajectory, kept for the TCC
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
      if (s == 0 || last < best_fde || (last != last && best_fde == best_fde)) {
        best_fde = last;
        best_idx = s;
        if (TCC) {
#pragma unroll
          for (int q = 0; q < 2 * T / 4; ++q) best_row[q] = v[q];
        }
      }
      if (++slot == NSTAGE) { slot = 0; parity ^= 1u; }
    }
    if (live) {
      ade[i] = best_ade;
      fde[i] = best_fde;
      if (argmin_fde) argmin_fde[i] = best_idx;
      if (TCC) {
        float pb[2 * T];
#pragma unroll
        for (int q = 0; q < 2 * T / 4; ++q) {
          pb[4 * q] = best_row[q].x; pb[4 * q + 1] = best_row[q].y; pb[4 * q + 2] = best_row[q].z; pb[4 * q + 3] = best_row[q].w;
        }
        tcc[i] = tcc_of<T>(pb, g);
      }
    }
  }
}

// Any T: one thread per pedestrian, direct global reads.
__global__ void ade_fde_generic(const float* __restrict__ pred, const float* __restrict__ gt, int s_total, int64_t n,
                                int t, float* __restrict__ ade, float* __restrict__ fde,
                                int32_t* __restrict__ argmin_fde, float* __restrict__ tcc) {
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
  if (tcc) {   // re-read the best sample (generic path only; the fast path keeps it in registers)
    const float2* p = reinterpret_cast<const float2*>(pred) + ((int64_t)best_idx * n + i) * t;
    const float factor = 1.0f / (float)(t - 1);
    float out = 0.f;
    for (int c = 0; c < 2; ++c) {
      float mp = 0.f, mg = 0.f;
      for (int q = 0; q < t; ++q) {
        const float2 a = __ldg(p + q), b = __ldg(g + q);
        mp += c ? a.y : a.x;
        mg += c ? b.y : b.x;
      }
      mp /= (float)t;
      mg /= (float)t;
      float cpg = 0.f, cpp = 0.f, cgg = 0.f;
      for (int q = 0; q < t; ++q) {
        const float2 a = __ldg(p + q), b = __ldg(g + q);
        const float dp = (c ? a.y : a.x) - mp, dg = (c ? b.y : b.x) - mg;
        cpg = fmaf(factor * dp, dg, cpg);
        cpp = fmaf(factor * dp, dp, cpp);
        cgg = fmaf(factor * dg, dg, cgg);
      }
      const float raw = cpg / sqrtf(cpp) / sqrtf(cgg);
      out += (raw == raw) ? fminf(fmaxf(raw, -1.f), 1.f) : 0.f;
    }
    tcc[i] = out / 2.0f;
  }
}

// Collision rate (utils/metrics.py:133-155).  One thread per pedestrian j; for every sample the 14 interpolated
// positions (first frame + cumulative quarter steps, summed sequentially in fp32 as torch.cumsum does on the CPU) of a
// tile of pedestrians i are staged in shared memory and compared against j's.  out[j] = count / S * 100.
constexpr int COL_STEPS = 14, COL_THREADS = 128;

__device__ __forceinline__ void col_dense_positions(const float2* __restrict__ row, int t, float2 (&d)[COL_STEPS]) {
  // dense step m = frame m/4 + (m%4) quarter steps; torch builds it as a running sum of rel/4 increments
  float2 cur = __ldg(row);
  d[0] = cur;
  int m = 1;
  for (int f = 0; f + 1 < t && m < COL_STEPS; ++f) {
    const float2 a = __ldg(row + f), b = __ldg(row + f + 1);
    const float ix = __fdiv_rn(__fsub_rn(b.x, a.x), 4.0f), iy = __fdiv_rn(__fsub_rn(b.y, a.y), 4.0f);
    for (int q = 0; q < 4 && m < COL_STEPS; ++q, ++m) {
      cur.x = __fadd_rn(cur.x, ix);
      cur.y = __fadd_rn(cur.y, iy);
      d[m] = cur;
    }
  }
  for (; m < COL_STEPS; ++m) d[m] = cur;   // t too short for 14 dense steps: the reference slices what exists
}

__global__ void __launch_bounds__(COL_THREADS) col_kernel(const float* __restrict__ pred, int s_total, int64_t n, int t,
                                                          int steps, float thres, float* __restrict__ out) {
  __shared__ float2 tile[COL_THREADS][COL_STEPS + 1];
  const int64_t j = (int64_t)blockIdx.x * COL_THREADS + threadIdx.x;
  int count = 0;
  for (int s = 0; s < s_total; ++s) {
    const float2* base = reinterpret_cast<const float2*>(pred) + (int64_t)s * n * t;
    float2 mine[COL_STEPS];
    if (j < n) col_dense_positions(base + j * t, t, mine);
    bool hit = false;
    for (int64_t i0 = 0; i0 < n; i0 += COL_THREADS) {
      __syncthreads();
      const int64_t i = i0 + threadIdx.x;
      if (i < n) {
        float2 d[COL_STEPS];
        col_dense_positions(base + i * t, t, d);
#pragma unroll
        for (int m = 0; m < COL_STEPS; ++m) tile[threadIdx.x][m] = d[m];
      }
      __syncthreads();
      if (j < n) {
        const int lim = (int)((n - i0) < COL_THREADS ? (n - i0) : COL_THREADS);
        for (int k = 0; k < lim; ++k) {
          if (i0 + k == j) continue;           // the reference adds 1 on the diagonal: never below the threshold
          float best = INFINITY;
          for (int m = 0; m < steps; ++m) {
            const float2 o = tile[k][m];
            const float dx = __fsub_rn(mine[m].x, o.x), dy = __fsub_rn(mine[m].y, o.y);
            best = fminf(best, sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
          }
          hit |= best < thres;
        }
      }
    }
    count += hit ? 1 : 0;
  }
  if (j < n) out[j] = __fmul_rn(__fdiv_rn((float)count, (float)s_total), 100.0f);
}

template <int WARPS, int NSTAGE, int MINB = 1>
static int launch_ade(const float* pred, const float* gt, int s, int64_t n, int64_t n_tiles, float* ade, float* fde,
                      int32_t* argmin_fde, float* tcc, int blocks_per_sm, cudaStream_t st) {
  using L = AdeSmem<12, NSTAGE, WARPS>;
  auto kern = tcc ? ade_fde_fast<12, NSTAGE, WARPS, MINB, true> : ade_fde_fast<12, NSTAGE, WARPS, MINB, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "ade_fde_fast: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int64_t grid = (n_tiles + WARPS - 1) / WARPS;
  // blocks_per_sm > 0: persistent grid of that many blocks per SM; 0: one tile per warp, the hardware block scheduler
  // balances the load (best when a tile is a sizeable unit of work, as here: S slabs of 3 KB)
  const int64_t cap = blocks_per_sm > 0 ? (int64_t)sm_count() * blocks_per_sm : ((int64_t)1 << 30);
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, WARPS * 32, L::bytes, st>>>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, tcc);
  return check_launch("ade_fde_fast");
}

}  // namespace et

using namespace et;

extern "C" int et_col(const float* pred, int s, int64_t n, int t, float thres, float* col, et_stream_t stream) {
  ET_REQUIRE(n >= 0 && s >= 1 && t >= 1, ET_ERR_BADARG, "et_col: bad shape (S=%d, N=%lld, T=%d)", s, (long long)n, t);
  if (n == 0) return ET_OK;
  ET_REQUIRE((pred && col) || n == 0, ET_ERR_BADARG, "et_col: null pointer");
  const int avail = 4 * (t - 1) + 1;                 // dense steps that exist
  const int steps = avail < COL_STEPS ? avail : COL_STEPS;
  col_kernel<<<(unsigned)((n + COL_THREADS - 1) / COL_THREADS), COL_THREADS, 0, as_stream(stream)>>>(pred, s, n, t, steps, thres, col);
  return check_launch("col_kernel");
}

extern "C" int et_ade_fde(const float* pred, const float* gt, int s, int64_t n, int t, float* ade, float* fde,
                          int32_t* argmin_fde, float* tcc, et_stream_t stream) {
  ET_REQUIRE(n >= 0 && s >= 1 && t >= 1, ET_ERR_BADARG, "et_ade_fde: bad shape (S=%d, N=%lld, T=%d)", s, (long long)n, t);
  if (n == 0) return ET_OK;
  ET_REQUIRE((pred && gt && ade && fde) || n == 0, ET_ERR_BADARG, "et_ade_fde: null pointer");
  ET_REQUIRE(aligned16(pred) && aligned16(gt), ET_ERR_ALIGN, "et_ade_fde: pred / gt must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (t == 12 && n >= 32) {
    const int64_t n_tiles = (n + 31) / 32;
    // <warps per block, ring depth, min blocks/SM>(..., blocks per SM of a persistent grid | 0 = one tile per warp).
    // Measured at S = 20, N = 2e5 on B200: 79 us for the default vs 87-93 us for the persistent shapes.
    switch (tune_get(ET_TUNE_ADE_CONFIG)) {
      case 1: return launch_ade<8, 4, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, tcc, 2, st);
      case 2: return launch_ade<16, 2, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, tcc, 2, st);
      case 3: return launch_ade<8, 3, 3>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, tcc, 0, st);
      default: return launch_ade<16, 2, 1>(pred, gt, s, n, n_tiles, ade, fde, argmin_fde, tcc, 0, st);
    }
  }
  ade_fde_generic<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pred, gt, s, n, t, ade, fde, argmin_fde, tcc);
  return check_launch("ade_fde_generic");
}
