// Descriptor path: TrajNorm, ETDescriptor.projection / reconstruction and the fused
// project->reconstruct round trip (reference: EigenTrajectory/normalizer.py,
// EigenTrajectory/descriptor.py, EigenTrajectory/anchor.py:76-88).
//
// Data layout in HBM (all fp32, exactly the reference's tensors):
//   trajectories (N,T,2) row = one pedestrian, 2T contiguous floats (64 B obs / 96 B pred)
//   coefficients (k,N) or (k,N,S), k-major
//   reconstructions (S,N,T,2)
// Every kernel is one pass over its inputs; the roofline that bounds each one is HBM.
#include "et_common.cuh"
#include "et_tma.cuh"

namespace et {

constexpr int UPITCH = 8;  // floats per row of a basis staged in shared memory (k <= 8 fast path)

// Stage U (rows, k) row-major into shared memory with row pitch UPITCH, zero padded.
template <int ROWS, int K>
__device__ __forceinline__ void stage_basis(float* Us, const float* __restrict__ U, int tid, int nthreads) {
  for (int e = tid; e < ROWS * UPITCH; e += nthreads) {
    const int r = e / UPITCH, j = e % UPITCH;
    Us[e] = (j < K) ? __ldg(U + r * K + j) : 0.f;
  }
}

// c[j] = sum_i U[i][j] x[i], ascending i, one FMA chain per coefficient.
template <int T2, int K>
__device__ __forceinline__ void project_row(const float (&x)[T2], const float* Us, float (&c)[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) c[j] = 0.f;
#pragma unroll
  for (int i = 0; i < T2; ++i) {
    const float4 u0 = *reinterpret_cast<const float4*>(Us + i * UPITCH);
    const float4 u1 = *reinterpret_cast<const float4*>(Us + i * UPITCH + 4);
    const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int j = 0; j < K; ++j) c[j] = fmaf(u[j], x[i], c[j]);
  }
}

// y[i] = sum_j U[i][j] c[j], ascending j.
template <int T2, int K>
__device__ __forceinline__ void unproject_row(const float (&c)[K], const float* Us, float (&y)[T2]) {
  // compiler barrier: re-read the basis from shared memory instead of keeping the 2T*k values that
  // project_row just loaded alive in registers (that costs ~140 registers and spills)
  asm volatile("" ::: "memory");
#pragma unroll
  for (int i = 0; i < T2; ++i) {
    const float4 u0 = *reinterpret_cast<const float4*>(Us + i * UPITCH);
    const float4 u1 = *reinterpret_cast<const float4*>(Us + i * UPITCH + 4);
    const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fmaf(u[j], c[j], acc);
    y[i] = acc;
  }
}

template <int T2>
__device__ __forceinline__ void normalize_row(float (&x)[T2], const NormState& st, int flags) {
#pragma unroll
  for (int t = 0; t < T2 / 2; ++t) norm_fwd(x[2 * t], x[2 * t + 1], st, flags);
}
template <int T2>
__device__ __forceinline__ void normalize_row(float (&x)[T2], const AffineFwd& af) {
#pragma unroll
  for (int t = 0; t < T2 / 2; ++t) affine_fwd(x[2 * t], x[2 * t + 1], af);
}
template <int T2>
__device__ __forceinline__ void denormalize_row(float (&x)[T2], const AffineBwd& af) {
#pragma unroll
  for (int t = 0; t < T2 / 2; ++t) affine_bwd(x[2 * t], x[2 * t + 1], af);
}

template <int T2>
__device__ __forceinline__ void load_row_global(float (&x)[T2], const float* __restrict__ base, int64_t i) {
  const float4* p = reinterpret_cast<const float4*>(base + i * T2);
#pragma unroll
  for (int c = 0; c < T2 / 4; ++c) {
    const float4 v = __ldg(p + c);
    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
}
template <int T2>
__device__ __forceinline__ void store_row_global(const float (&x)[T2], float* __restrict__ base, int64_t i) {
  float4* p = reinterpret_cast<float4*>(base + i * T2);
#pragma unroll
  for (int c = 0; c < T2 / 4; ++c) p[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
}

// =======================================================================================
// 1. Elementwise normaliser kernels (any T): one thread per (pedestrian, frame) point.
// =======================================================================================
__global__ void norm_params_kernel(const float* __restrict__ obs, int64_t n, int t_obs, int flags,
                                   float* __restrict__ ori, float* __restrict__ rot, float* __restrict__ sca) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2* row = reinterpret_cast<const float2*>(obs) + i * t_obs;
  const float2 last = __ldg(row + t_obs - 1), third = __ldg(row + t_obs - 3);
  const NormState st = make_norm_state(last.x, last.y, third.x, third.y);
  store_norm_state(ori, rot, sca, i, st, flags);
}

template <bool FWD>
__global__ void normalize_points_kernel(const float* __restrict__ traj, int64_t n, int t, int flags,
                                        const float* __restrict__ ori, const float* __restrict__ rot,
                                        const float* __restrict__ sca, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * t) return;
  const int64_t i = e / t;
  const NormState st = load_norm_state(ori, rot, sca, i, flags);
  float2 p = __ldg(reinterpret_cast<const float2*>(traj) + e);
  if (FWD) {
    norm_fwd(p.x, p.y, st, flags);
  } else {
    // the reference divides by sca (normalizer.py:56); keep the true division here
    if (flags & ET_NORM_SCA) { p.x = p.x / st.sca; p.y = p.y / st.sca; }
    norm_bwd(p.x, p.y, st, 1.f, flags & ~ET_NORM_SCA);
  }
  reinterpret_cast<float2*>(out)[e] = p;
}

// =======================================================================================
// 2. Generic (any T <= 32, k <= 32) thread-per-pedestrian kernels.  Correct for every shape the
//    API admits; the (8, 12, 6) configuration of the reference takes the fast paths below.
// =======================================================================================
constexpr int GMAX2T = 2 * ET_MAX_T;

__device__ __forceinline__ void generic_load_basis(float* Us, const float* U, int rows, int k) {
  for (int e = threadIdx.x; e < rows * k; e += blockDim.x) Us[e] = __ldg(U + e);
}

__global__ void to_et_space_generic(const float* __restrict__ traj, int64_t n, int t2, const float* __restrict__ U,
                                    int k, float* __restrict__ C) {
  extern __shared__ float Us[];
  generic_load_basis(Us, U, t2, k);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[GMAX2T];
  for (int r = 0; r < t2; ++r) x[r] = __ldg(traj + i * t2 + r);
  for (int j = 0; j < k; ++j) {
    float acc = 0.f;
    for (int r = 0; r < t2; ++r) acc = fmaf(Us[r * k + j], x[r], acc);
    C[(int64_t)j * n + i] = acc;
  }
}

__global__ void to_euclidean_generic(const float* __restrict__ C, int64_t ldc_k, int64_t ldc_n, int64_t n, int t2,
                                     const float* __restrict__ U, int k, float* __restrict__ traj) {
  extern __shared__ float Us[];
  generic_load_basis(Us, U, t2, k);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c[ET_MAX_K];
  for (int j = 0; j < k; ++j) c[j] = __ldg(C + j * ldc_k + i * ldc_n);
  for (int r = 0; r < t2; ++r) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc = fmaf(Us[r * k + j], c[j], acc);
    traj[i * t2 + r] = acc;
  }
}

// projection (+ optional round trip) for arbitrary shapes
__global__ void project_generic(const float* __restrict__ obs, const float* __restrict__ pred, int64_t n, int to2,
                                int tp2, const float* __restrict__ U_obs, const float* __restrict__ U_pred, int k,
                                int flags, float* __restrict__ C_obs, float* __restrict__ C_pred,
                                float* __restrict__ ori, float* __restrict__ rot, float* __restrict__ sca,
                                float* __restrict__ rec_obs, float* __restrict__ rec_pred) {
  extern __shared__ float Us[];
  float* Uo = Us;
  float* Up = Us + to2 * k;
  generic_load_basis(Uo, U_obs, to2, k);
  if (pred) generic_load_basis(Up, U_pred, tp2, k);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[GMAX2T], c[ET_MAX_K];
  for (int r = 0; r < to2; ++r) x[r] = __ldg(obs + i * to2 + r);
  const NormState st = make_norm_state(x[to2 - 2], x[to2 - 1], x[to2 - 6], x[to2 - 5]);
  store_norm_state(ori, rot, sca, i, st, flags);
  const float inv_sca = 1.0f / st.sca;
  for (int pass = 0; pass < 2; ++pass) {
    const int t2 = pass ? tp2 : to2;
    const float* Ub = pass ? Up : Uo;
    float* Cout = pass ? C_pred : C_obs;
    float* rec = pass ? rec_pred : rec_obs;
    if (pass) {
      if (!pred) break;
      for (int r = 0; r < t2; ++r) x[r] = __ldg(pred + i * t2 + r);
    }
    for (int r = 0; r < t2; r += 2) norm_fwd(x[r], x[r + 1], st, flags);
    for (int j = 0; j < k; ++j) {
      float acc = 0.f;
      for (int r = 0; r < t2; ++r) acc = fmaf(Ub[r * k + j], x[r], acc);
      c[j] = acc;
      if (Cout) Cout[(int64_t)j * n + i] = acc;
    }
    if (rec) {
      for (int r = 0; r < t2; r += 2) {
        float a = 0.f, b = 0.f;
        for (int j = 0; j < k; ++j) {
          a = fmaf(Ub[r * k + j], c[j], a);
          b = fmaf(Ub[(r + 1) * k + j], c[j], b);
        }
        norm_bwd(a, b, st, inv_sca, flags);
        rec[i * t2 + r] = a;
        rec[i * t2 + r + 1] = b;
      }
    }
  }
}

// reconstruction for arbitrary shapes: one thread per (pedestrian, sample)
__global__ void reconstruct_generic(const float* __restrict__ C, const float* __restrict__ anchor, int64_t n, int s,
                                    int k, int t2, const float* __restrict__ U, int flags,
                                    const float* __restrict__ ori, const float* __restrict__ rot,
                                    const float* __restrict__ sca, float* __restrict__ out) {
  extern __shared__ float Us[];
  generic_load_basis(Us, U, t2, k);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * s) return;
  const int64_t i = e / s;
  const int si = (int)(e - i * s);
  const NormState st = load_norm_state(ori, rot, sca, i, flags);
  const float inv_sca = 1.0f / st.sca;
  float c[ET_MAX_K];
  for (int j = 0; j < k; ++j) {
    float v = __ldg(C + ((int64_t)j * n + i) * s + si);
    if (anchor) v = __ldg(anchor + j * s + si) + v;
    c[j] = v;
  }
  float* o = out + ((int64_t)si * n + i) * t2;
  for (int r = 0; r < t2; r += 2) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < k; ++j) {
      a = fmaf(Us[r * k + j], c[j], a);
      b = fmaf(Us[(r + 1) * k + j], c[j], b);
    }
    norm_bwd(a, b, st, inv_sca, flags);
    o[r] = a;
    o[r + 1] = b;
  }
}

// d out / d C for arbitrary shapes (see reconstruct_bwd fast path for the derivation)
__global__ void reconstruct_bwd_generic(const float* __restrict__ grad_out, int64_t n, int s, int k, int t2,
                                        const float* __restrict__ U, int flags, const float* __restrict__ rot,
                                        const float* __restrict__ sca, float* __restrict__ grad_C) {
  extern __shared__ float Us[];
  generic_load_basis(Us, U, t2, k);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * s) return;
  const int64_t i = e / s;
  const int si = (int)(e - i * s);
  NormState st = load_norm_state(nullptr, rot, sca, i, flags & ~ET_NORM_ORI);
  const float inv_sca = 1.0f / st.sca;
  float g[GMAX2T];
  const float* go = grad_out + ((int64_t)si * n + i) * t2;
  for (int r = 0; r < t2; r += 2) {
    float a = __ldg(go + r), b = __ldg(go + r + 1);
    if (flags & ET_NORM_ROT) {
      const float na = a * st.r00 + b * st.r10, nb = a * st.r01 + b * st.r11;
      a = na; b = nb;
    }
    if (flags & ET_NORM_SCA) { a *= inv_sca; b *= inv_sca; }
    g[r] = a; g[r + 1] = b;
  }
  for (int j = 0; j < k; ++j) {
    float acc = 0.f;
    for (int r = 0; r < t2; ++r) acc = fmaf(Us[r * k + j], g[r], acc);
    grad_C[((int64_t)j * n + i) * s + si] = acc;
  }
}

// =======================================================================================
// 3. Fast path (T_obs, T_pred, k) = (8, 12, 6): ETDescriptor.projection, thread per pedestrian.
//    236 algorithmic bytes per pedestrian (160 read, 48 + 28 written).
// =======================================================================================
template <int TO, int TP, int K>
__global__ void __launch_bounds__(128) project_fast(const float* __restrict__ obs, const float* __restrict__ pred,
                                                    int64_t n, const float* __restrict__ U_obs,
                                                    const float* __restrict__ U_pred, int flags,
                                                    float* __restrict__ C_obs, float* __restrict__ C_pred,
                                                    float* __restrict__ ori, float* __restrict__ rot,
                                                    float* __restrict__ sca) {
  __shared__ __align__(16) float Uo[2 * TO * UPITCH];
  __shared__ __align__(16) float Up[2 * TP * UPITCH];
  stage_basis<2 * TO, K>(Uo, U_obs, threadIdx.x, blockDim.x);
  if (pred) stage_basis<2 * TP, K>(Up, U_pred, threadIdx.x, blockDim.x);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xo[2 * TO];
  load_row_global<2 * TO>(xo, obs, i);
  const NormState st = make_norm_state(xo[2 * TO - 2], xo[2 * TO - 1], xo[2 * TO - 6], xo[2 * TO - 5]);
  store_norm_state(ori, rot, sca, i, st, flags);
  float c[K];
  const AffineFwd af = make_affine_fwd(st, flags);
  normalize_row<2 * TO>(xo, af);
  project_row<2 * TO, K>(xo, Uo, c);
#pragma unroll
  for (int j = 0; j < K; ++j) C_obs[(int64_t)j * n + i] = c[j];
  if (pred) {
    float xp[2 * TP];
    load_row_global<2 * TP>(xp, pred, i);
    normalize_row<2 * TP>(xp, af);
    project_row<2 * TP, K>(xp, Up, c);
#pragma unroll
    for (int j = 0; j < K; ++j) C_pred[(int64_t)j * n + i] = c[j];
  }
}

// =======================================================================================
// 4. Headline op, variant 1: direct global access, thread per pedestrian.
// =======================================================================================
template <int TO, int TP, int K>
__global__ void __launch_bounds__(128, 4) project_reconstruct_direct(
    const float* __restrict__ obs, const float* __restrict__ pred, int64_t n, const float* __restrict__ U_obs,
    const float* __restrict__ U_pred, int flags, float* __restrict__ rec_obs, float* __restrict__ rec_pred,
    float* __restrict__ C_obs, float* __restrict__ C_pred) {
  __shared__ __align__(16) float Uo[2 * TO * UPITCH];
  __shared__ __align__(16) float Up[2 * TP * UPITCH];
  stage_basis<2 * TO, K>(Uo, U_obs, threadIdx.x, blockDim.x);
  stage_basis<2 * TP, K>(Up, U_pred, threadIdx.x, blockDim.x);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xo[2 * TO], xp[2 * TP];
  load_row_global<2 * TO>(xo, obs, i);
  load_row_global<2 * TP>(xp, pred, i);
  const NormState st = make_norm_state(xo[2 * TO - 2], xo[2 * TO - 1], xo[2 * TO - 6], xo[2 * TO - 5]);
  const AffineFwd af = make_affine_fwd(st, flags);
  const AffineBwd ab = make_affine_bwd(st, flags);
  float c[K];
  {
    normalize_row<2 * TO>(xo, af);
    project_row<2 * TO, K>(xo, Uo, c);
    if (C_obs) {
#pragma unroll
      for (int j = 0; j < K; ++j) C_obs[(int64_t)j * n + i] = c[j];
    }
    unproject_row<2 * TO, K>(c, Uo, xo);
    denormalize_row<2 * TO>(xo, ab);
    store_row_global<2 * TO>(xo, rec_obs, i);
  }
  {
    normalize_row<2 * TP>(xp, af);
    project_row<2 * TP, K>(xp, Up, c);
    if (C_pred) {
#pragma unroll
      for (int j = 0; j < K; ++j) C_pred[(int64_t)j * n + i] = c[j];
    }
    unproject_row<2 * TP, K>(c, Up, xp);
    denormalize_row<2 * TP>(xp, ab);
    store_row_global<2 * TP>(xp, rec_pred, i);
  }
}

// =======================================================================================
// 5. Headline op, variant 2: persistent, warp-specialised, TMA-tiled (8, 12, 6).
//
//    Tile = 128 pedestrians = three boxes of one stage buffer:
//      [0, 8192)       obs  rows of 64 B, 64B-swizzled by the TMA engine
//      [8192, 16384)   pred columns 0..15  (64 B per row, 64B swizzle)
//      [16384, 20480)  pred columns 16..23 (32 B per row, 32B swizzle)
//    so that thread-per-row float4 reads and writes of shared memory are bank-conflict free.
//    Warp 4 (one lane) is the TMA producer: it loads tiles into a ring of NS stages and, when the
//    four consumer warps have overwritten a stage in place with the reconstructions, stores it.
//    No block-wide barrier in the steady state; consumers only ever wait on `full`.
// =======================================================================================
constexpr int PR_TILE = 128;
constexpr int PR_OBS_BYTES = PR_TILE * 64;
constexpr int PR_PA_BYTES = PR_TILE * 64;
constexpr int PR_PB_BYTES = PR_TILE * 32;
constexpr int PR_LOAD_BYTES = PR_OBS_BYTES + PR_PA_BYTES + PR_PB_BYTES;    // 20480: what one tile's TMA loads bring
constexpr int PR_C_BYTES = 6 * PR_TILE * 4;                                  // 3072: one (k, 128) coefficient box
template <bool TMAC>
struct PRStage { static constexpr int bytes = PR_LOAD_BYTES + (TMAC ? 2 * PR_C_BYTES : 0); };   // 20480 / 26624
constexpr int PR_CONSUMER_WARPS = PR_TILE / 32;
constexpr int PR_THREADS = (PR_CONSUMER_WARPS + 1) * 32;

struct PRMaps {
  CUtensorMap in_obs, in_pa, in_pb, out_obs, out_pa, out_pb, out_cobs, out_cpred;
};

template <int NS, bool TMAC>
constexpr size_t pr_smem_bytes() {
  return 1024 /* alignment slack */ + (size_t)NS * PRStage<TMAC>::bytes + (16 + 24) * UPITCH * 4 + 2 * NS * 8;
}

// TMAC: the coefficient rows of a tile are staged in shared memory as two (k, 128) boxes and written by the producer
// with tensor stores as well (needs N % 4 == 0 for the (k, N) tensor map); otherwise consumers store them directly.
// RECON: write the rank-k reconstructions back (the round trip); false = ETDescriptor.projection only, which instead
// writes the per-pedestrian normaliser state (ori / rot / sca).
template <int NS, int MINB, bool TMAC, bool RECON>
__global__ void __launch_bounds__(PR_THREADS, MINB) project_reconstruct_tma(const __grid_constant__ PRMaps maps, int64_t n,
                                                                       int n_tiles, const float* __restrict__ U_obs,
                                                                       const float* __restrict__ U_pred, int flags,
                                                                       float* __restrict__ C_obs,
                                                                       float* __restrict__ C_pred,
                                                                       float* __restrict__ ori, float* __restrict__ rot,
                                                                       float* __restrict__ sca) {
  constexpr int PR_STAGE_BYTES = PRStage<TMAC>::bytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* Uo = reinterpret_cast<float*>(stages + (size_t)NS * PR_STAGE_BYTES);
  float* Up = Uo + 16 * UPITCH;
  uint64_t* full = reinterpret_cast<uint64_t*>(Up + 24 * UPITCH);
  uint64_t* done = full + NS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], PR_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();   // barriers are live: the producer starts loading while the consumers stage the bases
  // Programmatic dependent launch: the launch that follows us in the stream may make its blocks resident (and run its
  // own prologue up to this point) as ours retire; nothing of OURS below may run before everything that precedes us in
  // the stream is complete and visible -- the bases, the trajectories and the output buffers are global memory.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (warp < PR_CONSUMER_WARPS) {
    stage_basis<16, 6>(Uo, U_obs, threadIdx.x, PR_CONSUMER_WARPS * 32);
    stage_basis<24, 6>(Up, U_pred, threadIdx.x, PR_CONSUMER_WARPS * 32);
    asm volatile("bar.sync 1, %0;" ::"n"(PR_CONSUMER_WARPS * 32) : "memory");   // consumers only
  }

  const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == PR_CONSUMER_WARPS) {
    // ------------------------------ producer ------------------------------
    if (lane == 0 && my_tiles > 0) {
      prefetch_tensormap(&maps.in_obs);  prefetch_tensormap(&maps.in_pa);  prefetch_tensormap(&maps.in_pb);
      if (RECON) { prefetch_tensormap(&maps.out_obs); prefetch_tensormap(&maps.out_pa); prefetch_tensormap(&maps.out_pb); }
      if (TMAC && C_obs) prefetch_tensormap(&maps.out_cobs);
      if (TMAC && C_pred) prefetch_tensormap(&maps.out_cpred);
      auto issue_load = [&](int it) {
        const int s = it % NS;
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * PR_TILE;
        uint8_t* st = stages + (size_t)s * PR_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], PR_LOAD_BYTES);
        tma_load_2d(st, &maps.in_obs, 0, row0, &full[s]);
        tma_load_2d(st + PR_OBS_BYTES, &maps.in_pa, 0, row0, &full[s]);
        tma_load_2d(st + PR_OBS_BYTES + PR_PA_BYTES, &maps.in_pb, 16, row0, &full[s]);
      };
      const int pre = my_tiles < NS ? my_tiles : NS;
      for (int it = 0; it < pre; ++it) issue_load(it);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % NS;
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * PR_TILE;
        uint8_t* st = stages + (size_t)s * PR_STAGE_BYTES;
        mbar_wait(&done[s], (uint32_t)((it / NS) & 1));
        if (RECON) {
          tma_store_2d(&maps.out_obs, 0, row0, st);
          tma_store_2d(&maps.out_pa, 0, row0, st + PR_OBS_BYTES);
          tma_store_2d(&maps.out_pb, 16, row0, st + PR_OBS_BYTES + PR_PA_BYTES);
        }
        if (TMAC) {
          if (C_obs) tma_store_2d(&maps.out_cobs, row0, 0, st + PR_LOAD_BYTES);
          if (C_pred) tma_store_2d(&maps.out_cpred, row0, 0, st + PR_LOAD_BYTES + PR_C_BYTES);
        }
        bulk_commit();
        // refill the stage whose store was issued one iteration ago
        if (it >= 1 && it - 1 + NS < my_tiles) {
          bulk_wait_read<1>();
          issue_load(it - 1 + NS);
        }
      }
      bulk_wait_all<0>();
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  const int r = warp * 32 + lane;
  const uint32_t sw64 = (uint32_t)((r >> 1) & 3) << 4;
  const uint32_t sw32 = (uint32_t)((r >> 2) & 1) << 4;
  for (int it = 0; it < my_tiles; ++it) {
    const int s = it % NS;
    const int64_t i = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PR_TILE + r;
    uint8_t* st = stages + (size_t)s * PR_STAGE_BYTES;
    uint8_t* so = st + r * 64;
    uint8_t* sa = st + PR_OBS_BYTES + r * 64;
    uint8_t* sb = st + PR_OBS_BYTES + PR_PA_BYTES + r * 32;
    mbar_wait(&full[s], (uint32_t)((it / NS) & 1));

    float co[6], cp[6];
    NormState nst;
    AffineFwd af;
    AffineBwd ab;
    {
      float xo[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(so + (((uint32_t)c << 4) ^ sw64));
        xo[4 * c] = v.x; xo[4 * c + 1] = v.y; xo[4 * c + 2] = v.z; xo[4 * c + 3] = v.w;
      }
      nst = make_norm_state(xo[14], xo[15], xo[10], xo[11]);
      af = make_affine_fwd(nst, flags);
      ab = make_affine_bwd(nst, flags);
      normalize_row<16>(xo, af);
      project_row<16, 6>(xo, Uo, co);
      if (RECON) {
        unproject_row<16, 6>(co, Uo, xo);
        denormalize_row<16>(xo, ab);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(so + (((uint32_t)c << 4) ^ sw64)) =
              make_float4(xo[4 * c], xo[4 * c + 1], xo[4 * c + 2], xo[4 * c + 3]);
      }
    }
    {
      float xp[24];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(sa + (((uint32_t)c << 4) ^ sw64));
        xp[4 * c] = v.x; xp[4 * c + 1] = v.y; xp[4 * c + 2] = v.z; xp[4 * c + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(sb + (((uint32_t)c << 4) ^ sw32));
        xp[16 + 4 * c] = v.x; xp[17 + 4 * c] = v.y; xp[18 + 4 * c] = v.z; xp[19 + 4 * c] = v.w;
      }
      normalize_row<24>(xp, af);
      project_row<24, 6>(xp, Up, cp);
      if (RECON) {
        unproject_row<24, 6>(cp, Up, xp);
        denormalize_row<24>(xp, ab);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<float4*>(sa + (((uint32_t)c << 4) ^ sw64)) =
              make_float4(xp[4 * c], xp[4 * c + 1], xp[4 * c + 2], xp[4 * c + 3]);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          *reinterpret_cast<float4*>(sb + (((uint32_t)c << 4) ^ sw32)) =
              make_float4(xp[16 + 4 * c], xp[17 + 4 * c], xp[18 + 4 * c], xp[19 + 4 * c]);
      }
    }
    if (!RECON && i < n) store_norm_state(ori, rot, sca, i, nst, flags);

    if (TMAC) {
      float* cb = reinterpret_cast<float*>(st + PR_LOAD_BYTES);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        cb[j * PR_TILE + r] = co[j];
        cb[6 * PR_TILE + j * PR_TILE + r] = cp[j];
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);

    if (!TMAC && i < n) {
      if (C_obs) {
#pragma unroll
        for (int j = 0; j < 6; ++j) C_obs[(int64_t)j * n + i] = co[j];
      }
      if (C_pred) {
#pragma unroll
        for (int j = 0; j < 6; ++j) C_pred[(int64_t)j * n + i] = cp[j];
      }
    }
  }
}

// =======================================================================================
// 6. ETDescriptor.reconstruction fused with ETAnchor.forward, (k, T, S) = (6, 12, 20).
//    2428 algorithmic bytes per pedestrian, 79 % of them the (S,N,T,2) output.
//    A block of four warps owns a tile of 32 pedestrians: six 1-D bulk loads bring the tile's (k, 32*S)
//    coefficient block into shared memory once; warp w reconstructs samples w, w+4, ... -- each lane one
//    pedestrian -- into a warp-private 32 x 96 B slab which one bulk store writes to out[s, n0:n0+32]
//    (contiguous 3 KB).  Two slabs per warp keep a store in flight while the next sample is computed.
//    The next tile's coefficient block and normaliser state are prefetched while the current tile is computed.
//    ~57 KB of shared memory per block => four blocks (16 warps) per SM.
// =======================================================================================
constexpr int REC_WARPS = 4;

template <int K, int T, int S, int CBUF = 1>
struct RecSmem {
  static constexpr int C_FLOATS = K * 32 * S;
  static constexpr int SLAB_FLOATS = 32 * 2 * T;
  static constexpr int U_FLOATS = 2 * T * UPITCH;
  static constexpr int A_FLOATS = K * S;
  static constexpr size_t bytes =
      128 + (size_t)(CBUF * C_FLOATS + REC_WARPS * 2 * SLAB_FLOATS + U_FLOATS + A_FLOATS) * 4 + (2 * REC_WARPS + 2) * 8;
};

template <int K, int T, int S>
__global__ void __launch_bounds__(REC_WARPS * 32) reconstruct_fast(const float* __restrict__ C,
                                                                   const float* __restrict__ anchor, int64_t n,
                                                                   int64_t n_tiles, const float* __restrict__ U,
                                                                   int flags, const float* __restrict__ ori,
                                                                   const float* __restrict__ rot,
                                                                   const float* __restrict__ sca,
                                                                   float* __restrict__ out) {
  using L = RecSmem<K, T, S, 2>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* Cbuf = reinterpret_cast<float*>(base);                 // two tiles: the next one loads while this one computes
  float* slabs = Cbuf + 2 * L::C_FLOATS;
  float* Us = slabs + REC_WARPS * 2 * L::SLAB_FLOATS;
  float* As = Us + L::U_FLOATS;
  uint64_t* full = reinterpret_cast<uint64_t*>(As + L::A_FLOATS);   // full[0], full[1]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_basis<2 * T, K>(Us, U, threadIdx.x, REC_WARPS * 32);
  for (int e = threadIdx.x; e < K * S; e += REC_WARPS * 32) As[e] = anchor ? __ldg(anchor + e) : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto load_tile = [&](int64_t tile, int buf) {     // thread 0 only
    const int64_t n0 = tile * 32;
    const int rows = (int)((n - n0) < 32 ? (n - n0) : 32);
    mbar_arrive_expect_tx(&full[buf], (uint32_t)(K * rows * S * 4));
#pragma unroll
    for (int j = 0; j < K; ++j)
      bulk_load(Cbuf + buf * L::C_FLOATS + j * 32 * S, C + ((int64_t)j * n + n0) * S, (uint32_t)(rows * S * 4), &full[buf]);
  };
  auto load_state = [&](int64_t tile) {
    const int64_t i = tile * 32 + lane;
    return load_norm_state(ori, rot, sca, i < n ? i : n - 1, flags);
  };

  float* slab = slabs + warp * 2 * L::SLAB_FLOATS;
  int slab_sel = 0;
  int64_t tile = blockIdx.x;
  if (tile >= n_tiles) return;
  if (threadIdx.x == 0) load_tile(tile, 0);
  NormState st_next = load_state(tile);
  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t n0 = tile * 32;
    const int rows = (int)((n - n0) < 32 ? (n - n0) : 32);
    const int64_t next = tile + gridDim.x;
    // prefetch the next tile: its coefficient block into the other buffer (every warp left it at the end of the previous
    // iteration) and its normaliser state into registers
    if (threadIdx.x == 0 && next < n_tiles) load_tile(next, buf ^ 1);
    const AffineBwd af = make_affine_bwd(st_next, flags);
    if (next < n_tiles) st_next = load_state(next);
    const float* Cs = Cbuf + buf * L::C_FLOATS;
    mbar_wait(&full[buf], (uint32_t)((it >> 1) & 1));
    for (int s = warp; s < S; s += REC_WARPS) {
      float c[K];
#pragma unroll
      for (int j = 0; j < K; ++j) c[j] = As[j * S + s] + Cs[j * 32 * S + lane * S + s];
      float y[2 * T];
      unproject_row<2 * T, K>(c, Us, y);
      denormalize_row<2 * T>(y, af);
      // the slab about to be overwritten was handed to a bulk store two samples ago
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      float* sl = slab + slab_sel * L::SLAB_FLOATS + lane * 2 * T;
#pragma unroll
      for (int q = 0; q < 2 * T / 4; ++q)
        reinterpret_cast<float4*>(sl)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store(out + ((int64_t)s * n + n0) * 2 * T, slab + slab_sel * L::SLAB_FLOATS, (uint32_t)(rows * 2 * T * 4));
        bulk_commit();
      }
      slab_sel ^= 1;
    }
    __syncthreads();   // every warp has finished reading this buffer before the tile after next is loaded into it
  }
  if (lane == 0) bulk_wait_all<0>();
}

// Gradient wrt C.  out[s,n,t,:] = (m_t / sca) R^T + ori with m = U c, so for an upstream gradient g:
// d m_t = (g_t R) / sca and d c = U^T d m.  Mirror image of reconstruct_fast: each warp bulk-loads the
// (32 x 96 B) gradient slabs of its samples (double-buffered), the block assembles the (k, 32*S) tile in
// shared memory and bulk-stores it once per k.
template <int K, int T, int S>
__global__ void __launch_bounds__(REC_WARPS * 32) reconstruct_bwd_fast(const float* __restrict__ grad_out, int64_t n,
                                                                       int64_t n_tiles, const float* __restrict__ U,
                                                                       int flags, const float* __restrict__ rot,
                                                                       const float* __restrict__ sca,
                                                                       float* __restrict__ grad_C) {
  using L = RecSmem<K, T, S>;
  static_assert((S / REC_WARPS) * REC_WARPS == S, "samples must split evenly over the warps");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float* Cs = reinterpret_cast<float*>(base);
  float* slabs = Cs + L::C_FLOATS;
  float* Us = slabs + REC_WARPS * 2 * L::SLAB_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Us + L::U_FLOATS + L::A_FLOATS) + 2;   // two per warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_basis<2 * T, K>(Us, U, threadIdx.x, REC_WARPS * 32);
  if (threadIdx.x < 2 * REC_WARPS) mbar_init(&bars[threadIdx.x], 1);
  if (threadIdx.x == 0) fence_barrier_init();
  __syncthreads();

  float* slab = slabs + warp * 2 * L::SLAB_FLOATS;
  uint64_t* bar = &bars[2 * warp];
  constexpr int PER_WARP = S / REC_WARPS;
  uint32_t cnt = 0;   // slabs consumed so far by this warp: buffer b = cnt & 1 is on its (cnt >> 1)-th phase
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * 32;
    const int rows = (int)((n - n0) < 32 ? (n - n0) : 32);
    const uint32_t slab_bytes = (uint32_t)(rows * 2 * T * 4);
    const int64_t i = n0 + lane;
    const NormState st = load_norm_state(nullptr, rot, sca, i < n ? i : n - 1, flags & ~ET_NORM_ORI);
    // d m = (g R) / sca as one 2x2 map per pedestrian
    const float inv_sca = (flags & ET_NORM_SCA) ? 1.0f / st.sca : 1.0f;
    const bool rotate = (flags & ET_NORM_ROT) != 0;
    const float q00 = (rotate ? st.r00 : 1.f) * inv_sca, q10 = (rotate ? st.r10 : 0.f) * inv_sca;
    const float q01 = (rotate ? st.r01 : 0.f) * inv_sca, q11 = (rotate ? st.r11 : 1.f) * inv_sca;
    if (lane == 0) {
      const uint32_t b = cnt & 1u;
      mbar_arrive_expect_tx(&bar[b], slab_bytes);
      bulk_load(slab + b * L::SLAB_FLOATS, grad_out + ((int64_t)warp * n + n0) * 2 * T, slab_bytes, &bar[b]);
    }
    if (threadIdx.x == 0) bulk_wait_read<0>();   // the previous tile's Cs stores have left shared memory
    __syncthreads();
    for (int q = 0; q < PER_WARP; ++q) {
      const int s = warp + q * REC_WARPS;
      const uint32_t b = cnt & 1u;
      if (lane == 0 && q + 1 < PER_WARP) {
        mbar_arrive_expect_tx(&bar[b ^ 1u], slab_bytes);
        bulk_load(slab + (b ^ 1u) * L::SLAB_FLOATS, grad_out + ((int64_t)(s + REC_WARPS) * n + n0) * 2 * T, slab_bytes,
                  &bar[b ^ 1u]);
      }
      mbar_wait(&bar[b], (cnt >> 1) & 1u);
      ++cnt;
      float g[2 * T];
      const float* sl = slab + b * L::SLAB_FLOATS + lane * 2 * T;
#pragma unroll
      for (int v4 = 0; v4 < 2 * T / 4; ++v4) {
        const float4 v = reinterpret_cast<const float4*>(sl)[v4];
        g[4 * v4] = v.x; g[4 * v4 + 1] = v.y; g[4 * v4 + 2] = v.z; g[4 * v4 + 3] = v.w;
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const float a = g[2 * t], bb = g[2 * t + 1];
        g[2 * t] = fmaf(a, q00, bb * q10);
        g[2 * t + 1] = fmaf(a, q01, bb * q11);
      }
      float c[K];
      project_row<2 * T, K>(g, Us, c);
#pragma unroll
      for (int j = 0; j < K; ++j) Cs[j * 32 * S + lane * S + s] = c[j];
      __syncwarp();   // all lanes finished reading slab b before it is refilled
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < K; ++j)
        bulk_store(grad_C + ((int64_t)j * n + n0) * S, Cs + j * 32 * S, (uint32_t)(rows * S * 4));
      bulk_commit();
    }
  }
  if (threadIdx.x == 0) bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------
static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static int check_shape(int64_t n, int t, int k) {
  if (n < 0) return fail(ET_ERR_BADARG, "n = %lld < 0", (long long)n);
  if (t < 1 || t > ET_MAX_T) return fail(ET_ERR_UNSUPPORTED, "T = %d outside [1, %d]", t, ET_MAX_T);
  if (k < 1 || k > ET_MAX_K) return fail(ET_ERR_UNSUPPORTED, "k = %d outside [1, %d]", k, ET_MAX_K);
  return ET_OK;
}

template <typename F>
static int ensure_smem(F kernel, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "%s: cudaFuncSetAttribute(%zu B): %s", what, bytes, cudaGetErrorString(e));
  return ET_OK;
}

template <int NS, int MINB, bool TMAC, bool RECON>
static int launch_pr_tma_c(const float* obs, const float* pred, int64_t n, const float* U_obs, const float* U_pred,
                           int flags, float* rec_obs, float* rec_pred, float* C_obs, float* C_pred, float* ori,
                           float* rot, float* sca, cudaStream_t stream) {
  PRMaps maps;
  int rc;
  if ((rc = make_tensor_map_2d(&maps.in_obs, obs, (uint64_t)n, 16, 64, 16, PR_TILE, 64))) return rc;
  if ((rc = make_tensor_map_2d(&maps.in_pa, pred, (uint64_t)n, 24, 96, 16, PR_TILE, 64))) return rc;
  if ((rc = make_tensor_map_2d(&maps.in_pb, pred, (uint64_t)n, 24, 96, 8, PR_TILE, 32))) return rc;
  if (RECON) {
    if ((rc = make_tensor_map_2d(&maps.out_obs, rec_obs, (uint64_t)n, 16, 64, 16, PR_TILE, 64))) return rc;
    if ((rc = make_tensor_map_2d(&maps.out_pa, rec_pred, (uint64_t)n, 24, 96, 16, PR_TILE, 64))) return rc;
    if ((rc = make_tensor_map_2d(&maps.out_pb, rec_pred, (uint64_t)n, 24, 96, 8, PR_TILE, 32))) return rc;
  }
  if (TMAC) {   // (k, N) row-major coefficient matrices, box = 128 pedestrians x 6 rows
    if (C_obs && (rc = make_tensor_map_2d(&maps.out_cobs, C_obs, 6, (uint32_t)n, (uint64_t)n * 4, PR_TILE, 6, 0))) return rc;
    if (C_pred && (rc = make_tensor_map_2d(&maps.out_cpred, C_pred, 6, (uint32_t)n, (uint64_t)n * 4, PR_TILE, 6, 0))) return rc;
  }
  constexpr size_t smem = pr_smem_bytes<NS, TMAC>();
  auto kern = project_reconstruct_tma<NS, MINB, TMAC, RECON>;
  if ((rc = ensure_smem(kern, smem, "project_reconstruct_tma"))) return rc;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PR_THREADS, smem);
  if (per_sm < 1) per_sm = 1;
  const int n_tiles = (int)((n + PR_TILE - 1) / PR_TILE);
  int grid = sm_count() * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  // launched with programmatic stream serialization: back-to-back calls overlap their launch latency and prologue with
  // the previous call's tail (the kernel orders itself with griddepcontrol.wait); ET_TUNE_PDL = 1 switches it off
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(PR_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tune_get(ET_TUNE_PDL) == 1 ? 0 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps, n, n_tiles, U_obs, U_pred, flags, C_obs, C_pred, ori, rot, sca);
  if (le != cudaSuccess) return fail(ET_ERR_CUDA, "project_reconstruct_tma: launch: %s", cudaGetErrorString(le));
  return check_launch(RECON ? "project_reconstruct_tma" : "project_reconstruct_tma(project)");
}

template <int NS, int MINB>
static int launch_pr_tma(const float* obs, const float* pred, int64_t n, const float* U_obs, const float* U_pred,
                         int flags, float* rec_obs, float* rec_pred, float* C_obs, float* C_pred, bool tma_c,
                         cudaStream_t stream) {
  // tensor stores for the coefficients need a 16-byte row pitch (N % 4 == 0) and aligned bases
  const bool ok = tma_c && (C_obs || C_pred) && n % 4 == 0 && aligned16(C_obs) && aligned16(C_pred);
  if (ok)
    return launch_pr_tma_c<NS, MINB, true, true>(obs, pred, n, U_obs, U_pred, flags, rec_obs, rec_pred, C_obs, C_pred, nullptr,
                                                 nullptr, nullptr, stream);
  return launch_pr_tma_c<NS, MINB, false, true>(obs, pred, n, U_obs, U_pred, flags, rec_obs, rec_pred, C_obs, C_pred, nullptr,
                                                nullptr, nullptr, stream);
}

// Grid of the reconstruction kernels: persistent (occupancy x SMs) by default; ET_TUNE_REC_BLOCKS_PER_SM > 0 overrides the
// blocks per SM, < 0 launches one block per tile.
template <typename F>
static int64_t rec_grid(F kern, size_t smem, int64_t n_tiles) {
  const int knob = tune_get(ET_TUNE_REC_BLOCKS_PER_SM);
  if (knob < 0) return n_tiles;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_WARPS * 32, smem);
  if (per_sm < 1) per_sm = 1;
  if (knob > 0 && knob < per_sm) per_sm = knob;
  int64_t grid = (int64_t)sm_count() * per_sm;
  return grid > n_tiles ? n_tiles : grid;
}

}  // namespace et

using namespace et;

extern "C" {

int et_norm_params(const float* obs, int64_t n, int t_obs, int flags, float* ori, float* rot, float* sca,
                   et_stream_t stream) {
  if (n == 0) return ET_OK;
  ET_REQUIRE(obs || n == 0, ET_ERR_BADARG, "et_norm_params: obs is null");
  ET_REQUIRE(t_obs >= 3 && t_obs <= ET_MAX_T, ET_ERR_UNSUPPORTED, "et_norm_params: T_obs = %d outside [3, %d]", t_obs, ET_MAX_T);
  ET_REQUIRE(n >= 0, ET_ERR_BADARG, "et_norm_params: n < 0");
  ET_REQUIRE(aligned16(obs) && aligned16(rot), ET_ERR_ALIGN, "et_norm_params: pointers must be 16-byte aligned");
  norm_params_kernel<<<blocks_for(n, 256), 256, 0, as_stream(stream)>>>(obs, n, t_obs, flags, ori, rot, sca);
  return check_launch("norm_params_kernel");
}

static int normalize_common(bool fwd, const float* traj, int64_t n, int t, int flags, const float* ori,
                            const float* rot, const float* sca, float* out, et_stream_t stream) {
  if (n == 0) return ET_OK;
  ET_REQUIRE((traj && out) || n == 0, ET_ERR_BADARG, "et_(de)normalize: null trajectory pointer");
  ET_REQUIRE(t >= 1 && n >= 0, ET_ERR_BADARG, "et_(de)normalize: bad shape");
  ET_REQUIRE(!(flags & ET_NORM_ORI) || ori, ET_ERR_BADARG, "et_(de)normalize: ori flag set but ori is null");
  ET_REQUIRE(!(flags & ET_NORM_ROT) || rot, ET_ERR_BADARG, "et_(de)normalize: rot flag set but rot is null");
  ET_REQUIRE(!(flags & ET_NORM_SCA) || sca, ET_ERR_BADARG, "et_(de)normalize: sca flag set but sca is null");
  ET_REQUIRE(aligned16(traj) && aligned16(out) && aligned16(rot), ET_ERR_ALIGN, "et_(de)normalize: alignment");
  if (fwd)
    normalize_points_kernel<true><<<blocks_for(n * t, 256), 256, 0, as_stream(stream)>>>(traj, n, t, flags, ori, rot, sca, out);
  else
    normalize_points_kernel<false><<<blocks_for(n * t, 256), 256, 0, as_stream(stream)>>>(traj, n, t, flags, ori, rot, sca, out);
  return check_launch("normalize_points_kernel");
}

int et_normalize(const float* traj, int64_t n, int t, int flags, const float* ori, const float* rot,
                 const float* sca, float* out, et_stream_t stream) {
  return normalize_common(true, traj, n, t, flags, ori, rot, sca, out, stream);
}

int et_denormalize(const float* traj, int64_t n, int t, int flags, const float* ori, const float* rot,
                   const float* sca, float* out, et_stream_t stream) {
  return normalize_common(false, traj, n, t, flags, ori, rot, sca, out, stream);
}

int et_to_et_space(const float* traj, int64_t n, int t, const float* U, int k, float* C, et_stream_t stream) {
  int rc = check_shape(n, t, k);
  if (rc) return rc;
  if (n == 0) return ET_OK;
  ET_REQUIRE((traj && U && C) || n == 0, ET_ERR_BADARG, "et_to_et_space: null pointer");
  to_et_space_generic<<<blocks_for(n, 128), 128, 2 * t * k * sizeof(float), as_stream(stream)>>>(traj, n, 2 * t, U, k, C);
  return check_launch("to_et_space_generic");
}

int et_to_euclidean_space(const float* C, int64_t ldc_k, int64_t ldc_n, int64_t n, int t, const float* U, int k,
                          float* traj, et_stream_t stream) {
  int rc = check_shape(n, t, k);
  if (rc) return rc;
  if (n == 0) return ET_OK;
  ET_REQUIRE((traj && U && C) || n == 0, ET_ERR_BADARG, "et_to_euclidean_space: null pointer");
  to_euclidean_generic<<<blocks_for(n, 128), 128, 2 * t * k * sizeof(float), as_stream(stream)>>>(C, ldc_k, ldc_n, n, 2 * t, U, k, traj);
  return check_launch("to_euclidean_generic");
}

int et_project(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, const float* U_obs,
               const float* U_pred, int k, int flags, float* C_obs, float* C_pred, float* ori, float* rot,
               float* sca, et_stream_t stream) {
  int rc = check_shape(n, t_obs, k);
  if (rc) return rc;
  if (pred && (rc = check_shape(n, t_pred, k))) return rc;
  ET_REQUIRE(t_obs >= 3, ET_ERR_UNSUPPORTED, "et_project: T_obs = %d < 3", t_obs);
  if (n == 0) return ET_OK;
  ET_REQUIRE((obs && U_obs && C_obs) || n == 0, ET_ERR_BADARG, "et_project: obs / U_obs / C_obs null");
  ET_REQUIRE(!pred || (U_pred && C_pred), ET_ERR_BADARG, "et_project: pred given but U_pred / C_pred null");
  ET_REQUIRE(aligned16(obs) && aligned16(pred) && aligned16(rot), ET_ERR_ALIGN, "et_project: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (t_obs == 8 && pred && t_pred == 12 && k == 6 && n >= 4096 && n < ((int64_t)1 << 31) && n % 4 == 0 && aligned16(C_obs) &&
      aligned16(C_pred) && aligned16(ori))
    return launch_pr_tma_c<4, 2, true, false>(obs, pred, n, U_obs, U_pred, flags, nullptr, nullptr, C_obs, C_pred, ori, rot, sca, st);
  if (t_obs == 8 && (t_pred == 12 || !pred) && k == 6) {
    project_fast<8, 12, 6><<<blocks_for(n, 128), 128, 0, st>>>(obs, pred, n, U_obs, U_pred, flags, C_obs, C_pred, ori, rot, sca);
    return check_launch("project_fast");
  }
  const size_t smem = (size_t)(2 * t_obs + (pred ? 2 * t_pred : 0)) * k * sizeof(float);
  project_generic<<<blocks_for(n, 128), 128, smem, st>>>(obs, pred, n, 2 * t_obs, 2 * t_pred, U_obs, U_pred, k, flags,
                                                          C_obs, C_pred, ori, rot, sca, nullptr, nullptr);
  return check_launch("project_generic");
}

int et_project_reconstruct(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred,
                           const float* U_obs, const float* U_pred, int k, int flags, float* rec_obs,
                           float* rec_pred, float* C_obs, float* C_pred, int variant, et_stream_t stream) {
  int rc = check_shape(n, t_obs, k);
  if (rc) return rc;
  if ((rc = check_shape(n, t_pred, k))) return rc;
  ET_REQUIRE(t_obs >= 3, ET_ERR_UNSUPPORTED, "et_project_reconstruct: T_obs = %d < 3", t_obs);
  if (n == 0) return ET_OK;
  ET_REQUIRE((obs && pred && U_obs && U_pred && rec_obs && rec_pred) || n == 0, ET_ERR_BADARG,
             "et_project_reconstruct: null pointer");
  ET_REQUIRE(aligned16(obs) && aligned16(pred) && aligned16(rec_obs) && aligned16(rec_pred), ET_ERR_ALIGN,
             "et_project_reconstruct: trajectory pointers must be 16-byte aligned");
  ET_REQUIRE(variant >= 0 && variant <= 4, ET_ERR_BADARG, "et_project_reconstruct: variant %d", variant);
  cudaStream_t st = as_stream(stream);
  const bool fast = (t_obs == 8 && t_pred == 12 && k == 6);
  if (!fast) {
    ET_REQUIRE(variant <= 1, ET_ERR_UNSUPPORTED, "et_project_reconstruct: TMA variants need (T_obs,T_pred,k) = (8,12,6)");
    const size_t smem = (size_t)(2 * t_obs + 2 * t_pred) * k * sizeof(float);
    project_generic<<<blocks_for(n, 128), 128, smem, st>>>(obs, pred, n, 2 * t_obs, 2 * t_pred, U_obs, U_pred, k,
                                                            flags, C_obs, C_pred, nullptr, nullptr, nullptr, rec_obs,
                                                            rec_pred);
    return check_launch("project_generic(round trip)");
  }
  if (variant == 0) variant = (n >= 4096 && n < (int64_t)1 << 31) ? 2 : 1;
  if (variant == 1) {
    project_reconstruct_direct<8, 12, 6><<<blocks_for(n, 128), 128, 0, st>>>(obs, pred, n, U_obs, U_pred, flags,
                                                                             rec_obs, rec_pred, C_obs, C_pred);
    return check_launch("project_reconstruct_direct");
  }
  ET_REQUIRE(n < (int64_t)1 << 31, ET_ERR_UNSUPPORTED, "et_project_reconstruct: TMA variants need n < 2^31");
  // 2: 4 stages, 2 blocks/SM, coefficients through TMA stores   3: 3 stages, 3 blocks/SM   4: 4 stages, 2 blocks/SM,
  // coefficients stored directly by the consumer warps
  if (variant == 2) return launch_pr_tma<4, 2>(obs, pred, n, U_obs, U_pred, flags, rec_obs, rec_pred, C_obs, C_pred, true, st);
  if (variant == 3) return launch_pr_tma<3, 3>(obs, pred, n, U_obs, U_pred, flags, rec_obs, rec_pred, C_obs, C_pred, false, st);
  return launch_pr_tma<4, 2>(obs, pred, n, U_obs, U_pred, flags, rec_obs, rec_pred, C_obs, C_pred, false, st);
}

int et_reconstruct(const float* C, const float* anchor, int64_t n, int s, int k, int t, const float* U, int flags,
                   const float* ori, const float* rot, const float* sca, float* out, et_stream_t stream) {
  int rc = check_shape(n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1, ET_ERR_BADARG, "et_reconstruct: S = %d", s);
  if (n == 0) return ET_OK;
  ET_REQUIRE((C && U && out) || n == 0, ET_ERR_BADARG, "et_reconstruct: null pointer");
  ET_REQUIRE(!(flags & ET_NORM_ORI) || ori, ET_ERR_BADARG, "et_reconstruct: ori flag set but ori is null");
  ET_REQUIRE(!(flags & ET_NORM_ROT) || rot, ET_ERR_BADARG, "et_reconstruct: rot flag set but rot is null");
  ET_REQUIRE(!(flags & ET_NORM_SCA) || sca, ET_ERR_BADARG, "et_reconstruct: sca flag set but sca is null");
  ET_REQUIRE(aligned16(C) && aligned16(out) && aligned16(rot), ET_ERR_ALIGN, "et_reconstruct: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (k == 6 && t == 12 && s == 20 && n >= 32) {
    using L = RecSmem<6, 12, 20, 2>;
    auto kern = reconstruct_fast<6, 12, 20>;
    if ((rc = ensure_smem(kern, L::bytes, "reconstruct_fast"))) return rc;
    const int64_t n_tiles = (n + 31) / 32;
    const int64_t grid = rec_grid(kern, L::bytes, n_tiles);
    kern<<<(unsigned)grid, REC_WARPS * 32, L::bytes, st>>>(C, anchor, n, n_tiles, U, flags, ori, rot, sca, out);
    return check_launch("reconstruct_fast");
  }
  reconstruct_generic<<<blocks_for(n * s, 128), 128, 2 * t * k * sizeof(float), st>>>(C, anchor, n, s, k, 2 * t, U,
                                                                                      flags, ori, rot, sca, out);
  return check_launch("reconstruct_generic");
}

int et_reconstruct_bwd(const float* grad_out, int64_t n, int s, int k, int t, const float* U, int flags,
                       const float* rot, const float* sca, float* grad_C, et_stream_t stream) {
  int rc = check_shape(n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1, ET_ERR_BADARG, "et_reconstruct_bwd: S = %d", s);
  if (n == 0) return ET_OK;
  ET_REQUIRE((grad_out && U && grad_C) || n == 0, ET_ERR_BADARG, "et_reconstruct_bwd: null pointer");
  ET_REQUIRE(!(flags & ET_NORM_ROT) || rot, ET_ERR_BADARG, "et_reconstruct_bwd: rot flag set but rot is null");
  ET_REQUIRE(!(flags & ET_NORM_SCA) || sca, ET_ERR_BADARG, "et_reconstruct_bwd: sca flag set but sca is null");
  ET_REQUIRE(aligned16(grad_out) && aligned16(grad_C) && aligned16(rot), ET_ERR_ALIGN, "et_reconstruct_bwd: alignment");
  cudaStream_t st = as_stream(stream);
  if (k == 6 && t == 12 && s == 20 && n >= 32) {
    using L = RecSmem<6, 12, 20>;
    auto kern = reconstruct_bwd_fast<6, 12, 20>;
    if ((rc = ensure_smem(kern, L::bytes, "reconstruct_bwd_fast"))) return rc;
    const int64_t n_tiles = (n + 31) / 32;
    const int64_t grid = rec_grid(kern, L::bytes, n_tiles);
    kern<<<(unsigned)grid, REC_WARPS * 32, L::bytes, st>>>(grad_out, n, n_tiles, U, flags, rot, sca, grad_C);
    return check_launch("reconstruct_bwd_fast");
  }
  reconstruct_bwd_generic<<<blocks_for(n * s, 128), 128, 2 * t * k * sizeof(float), st>>>(grad_out, n, s, k, 2 * t, U,
                                                                                          flags, rot, sca, grad_C);
  return check_launch("reconstruct_bwd_generic");
}

}  // extern "C"
