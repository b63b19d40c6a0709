// Eigen-basis construction (reference: ETDescriptor.truncated_SVD, EigenTrajectory/descriptor.py:91-114).
//
// The reference runs LAPACK gesdd on the wide (2T, N) view and discards Vt.  Here the (2T x 2T) Gram matrix
// G = M M^T is accumulated in ONE pass over the trajectories -- fp32 inputs, products and sums in float64
// on the FP64 tensor cores (DMMA m8n8k4; fp32 x fp32 products are exact in fp64) -- and a cyclic Jacobi
// eigen-solve on G yields U and S = sqrt(lambda).  A multi-GPU caller sums G over row shards (one
// all-reduce) between the two steps.  Small problems (N up to ~2000 rows) can instead run a one-sided
// Hestenes Jacobi directly on the shared-memory resident tall-skinny (N x 2T) matrix.
#include <math.h>
#include <stdlib.h>

#include "et_common.cuh"

namespace et {

// =======================================================================================
// 1. Gram pass, (T_obs, T_pred) = (8, 12): 160 algorithmic bytes and 436 fp64 FMA per pedestrian.
// =======================================================================================
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int GR_WARPS = 16;                 // one 16-warp block per SM: 148 partial records / barrier arrivals
constexpr int GR_PITCH = 40;                 // floats per staged pedestrian: 16 obs + 24 pred, conflict-free
constexpr int GR_NBLK_O = 3, GR_NBLK_P = 6;  // upper-triangular 8x8 blocks of the 16x16 / 24x24 Gram matrices
constexpr int GR_GO = 16 * 16, GR_GP = 24 * 24;
constexpr int GR_REC = (GR_NBLK_O + GR_NBLK_P) * 64;   // unique entries a warp accumulates: nine 8x8 blocks
constexpr size_t GR_SMEM = (size_t)GR_WARPS * 32 * GR_PITCH * sizeof(float);   // staging tiles; reused by the block fold
static_assert(GR_SMEM >= (size_t)GR_WARPS * GR_REC * sizeof(double), "block fold must fit the staging tiles");

// workspace layout: two uint32 barrier counters (zero on entry, zero again on exit), then at byte 128 the block
// partials, entry-major: GR_REC rows of gridDim.x doubles.  Cooperative launch (grid barrier before the fold).
//
// Each warp walks its tiles of 32 pedestrians with a one-tile register prefetch: the global loads of tile i+1 are in
// flight while the DMMA phase of tile i runs, so the fp64 tensor pipe is not left idle behind HBM latency.
template <int UNROLL>
__global__ void __launch_bounds__(GR_WARPS * 32) gram_fast(const float* __restrict__ obs, const float* __restrict__ pred,
                                                           int64_t n, int flags, double* __restrict__ G_obs,
                                                           double* __restrict__ G_pred, unsigned* __restrict__ ticket,
                                                           double* __restrict__ partials, float* __restrict__ pred_norm,
                                                           float* __restrict__ ori, float* __restrict__ rot,
                                                           float* __restrict__ sca) {
  extern __shared__ __align__(16) unsigned char gr_smem[];
  float* xs = reinterpret_cast<float*>(gr_smem);   // [GR_WARPS][32 * GR_PITCH]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  float* xw = xs + warp * (32 * GR_PITCH);

  double co[GR_NBLK_O][2], cp[GR_NBLK_P][2];
#pragma unroll
  for (int b = 0; b < GR_NBLK_O; ++b) co[b][0] = co[b][1] = 0.0;
#pragma unroll
  for (int b = 0; b < GR_NBLK_P; ++b) cp[b][0] = cp[b][1] = 0.0;

  const int64_t n_tiles = (n + 31) / 32;
  const int64_t wstride = (int64_t)gridDim.x * GR_WARPS;
  float4 raw[10];
  auto fetch = [&](int64_t tile) {
    const int64_t i = tile * 32 + lane;
    if (tile < n_tiles && i < n) {
      const float4* po = reinterpret_cast<const float4*>(obs + i * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c) raw[c] = __ldg(po + c);
      if (pred) {
        const float4* pp = reinterpret_cast<const float4*>(pred + i * 24);
#pragma unroll
        for (int c = 0; c < 6; ++c) raw[4 + c] = __ldg(pp + c);
      } else {
#pragma unroll
        for (int c = 4; c < 10; ++c) raw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 10; ++c) raw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // warp-major tile order: the n_tiles % (grid * GR_WARPS) left-over tiles land on the low warps of EVERY block, not on
  // all warps of the first blocks, so no SM (and no scheduler) carries more than one extra tile
  int64_t tile = (int64_t)warp * gridDim.x + blockIdx.x;
  fetch(tile);
  for (; tile < n_tiles; tile += wstride) {
    {
      float x[40];
#pragma unroll
      for (int c = 0; c < 10; ++c) {
        x[4 * c] = raw[c].x; x[4 * c + 1] = raw[c].y; x[4 * c + 2] = raw[c].z; x[4 * c + 3] = raw[c].w;
      }
      if (flags && tile * 32 + lane < n) {
        const NormState nst = make_norm_state(x[14], x[15], x[10], x[11]);
        const AffineFwd af = make_affine_fwd(nst, flags);
#pragma unroll
        for (int t = 0; t < 20; ++t) affine_fwd(x[2 * t], x[2 * t + 1], af);
        // parameter_initialization's other outputs from the same pass: normaliser state and the normalised futures
        if (ori || rot || sca) store_norm_state(ori, rot, sca, tile * 32 + lane, nst, flags);
      }
      if (pred_norm && tile * 32 + lane < n) {
        float4* po = reinterpret_cast<float4*>(pred_norm + (tile * 32 + lane) * 24);
#pragma unroll
        for (int c = 0; c < 6; ++c) stg_stream(po + c, make_float4(x[16 + 4 * c], x[17 + 4 * c], x[18 + 4 * c], x[19 + 4 * c]));
      }
      float4* row = reinterpret_cast<float4*>(xw + lane * GR_PITCH);
#pragma unroll
      for (int c = 0; c < 10; ++c) row[c] = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
    }
    fetch(tile + wstride);   // next tile's loads fly during the DMMA phase below
    __syncwarp();
    // 8 k-steps of 4 pedestrians; fragment f holds x[ped 4*ks + t4][8 f + g] (serves as A and as B)
#pragma unroll UNROLL
    for (int ks = 0; ks < 8; ++ks) {
      const float* xr = xw + (4 * ks + t4) * GR_PITCH + g;
      double f[5];
#pragma unroll
      for (int b = 0; b < 5; ++b) f[b] = (double)xr[8 * b];
      dmma_m8n8k4(co[0][0], co[0][1], f[0], f[0]);
      dmma_m8n8k4(co[1][0], co[1][1], f[0], f[1]);
      dmma_m8n8k4(co[2][0], co[2][1], f[1], f[1]);
      if (pred) {
        dmma_m8n8k4(cp[0][0], cp[0][1], f[2], f[2]);
        dmma_m8n8k4(cp[1][0], cp[1][1], f[2], f[3]);
        dmma_m8n8k4(cp[2][0], cp[2][1], f[2], f[4]);
        dmma_m8n8k4(cp[3][0], cp[3][1], f[3], f[3]);
        dmma_m8n8k4(cp[4][0], cp[4][1], f[3], f[4]);
        dmma_m8n8k4(cp[5][0], cp[5][1], f[4], f[4]);
      }
    }
    __syncwarp();
  }

  // ---- block fold (fixed warp order => deterministic): every warp parks its nine 8x8 blocks in the (now free)
  // staging area -- lane (g, t4) holds entries 8 g + 2 t4, +1 of each block, i.e. 16 contiguous bytes per lane --
  // and thread e sums entry e over the warps.  Partials go out entry-major so that the grid fold reads them coalesced.
  __syncthreads();
  double* wacc = reinterpret_cast<double*>(gr_smem);     // [GR_WARPS][GR_REC]
  {
    double2* mine = reinterpret_cast<double2*>(wacc + warp * GR_REC) + lane;
#pragma unroll
    for (int b = 0; b < GR_NBLK_O; ++b) mine[b * 32] = make_double2(co[b][0], co[b][1]);
#pragma unroll
    for (int b = 0; b < GR_NBLK_P; ++b) mine[(GR_NBLK_O + b) * 32] = make_double2(cp[b][0], cp[b][1]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < GR_REC; e += blockDim.x) {
    double sacc = 0.0;
#pragma unroll
    for (int w = 0; w < GR_WARPS; ++w) sacc += wacc[w * GR_REC + e];
    partials[(size_t)e * gridDim.x + blockIdx.x] = sacc;
  }
  // ---- grid fold: after the barrier the warps of the grid each sum one entry over all block partials; strictly
  // upper 8x8 blocks are mirrored into the lower triangle ----
  grid_barrier(ticket, gridDim.x);
  const int n_el = pred ? GR_REC : GR_NBLK_O * 64;
  for (int e = blockIdx.x * GR_WARPS + warp; e < n_el; e += gridDim.x * GR_WARPS) {
    const double tot = warp_fold_contig(partials + (size_t)e * gridDim.x, (int)gridDim.x, lane);
    if (lane == 0) {
      const int b = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
      if (b < GR_NBLK_O) {
        const int br = b == 2 ? 1 : 0, bc = b == 0 ? 0 : 1;                  // blocks (0,0), (0,1), (1,1)
        const int r = 8 * br + rr, c = 8 * bc + cc;
        G_obs[r * 16 + c] += tot;
        if (br != bc) G_obs[c * 16 + r] += tot;
      } else {
        const int q = b - GR_NBLK_O;                                         // (0,0), (0,1), (0,2), (1,1), (1,2), (2,2)
        const int br = q < 3 ? 0 : (q < 5 ? 1 : 2), bc = q < 3 ? q : (q < 5 ? q - 2 : 2);
        const int r = 8 * br + rr, c = 8 * bc + cc;
        G_pred[r * 24 + c] += tot;
        if (br != bc) G_pred[c * 24 + r] += tot;
      }
    }
  }
}

// Any (T_obs, T_pred): block-wide fp64 accumulation over chunks of 64 staged pedestrians; simple and slow,
// used off the fast path only.
constexpr int GG_THREADS = 512, GG_CHUNK = 64, GG_ACC = 16;

__global__ void __launch_bounds__(GG_THREADS) gram_generic(const float* __restrict__ obs, const float* __restrict__ pred,
                                                           int64_t n, int to2, int tp2, int flags,
                                                           double* __restrict__ G_obs, double* __restrict__ G_pred) {
  extern __shared__ float xs[];   // GG_CHUNK rows of (to2 + tp2) floats
  const int pitch = to2 + tp2;
  const int n_out = to2 * to2 + tp2 * tp2;
  double acc[GG_ACC];
  for (int q = 0; q < GG_ACC; ++q) acc[q] = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * GG_CHUNK; base < n; base += (int64_t)gridDim.x * GG_CHUNK) {
    for (int e = threadIdx.x; e < GG_CHUNK * pitch; e += GG_THREADS) {
      const int m = e / pitch, r = e % pitch;
      const int64_t i = base + m;
      float v = 0.f;
      if (i < n) v = (r < to2) ? __ldg(obs + i * to2 + r) : __ldg(pred + i * tp2 + (r - to2));
      xs[e] = v;
    }
    __syncthreads();
    if (flags && threadIdx.x < GG_CHUNK && base + threadIdx.x < n) {
      float* row = xs + threadIdx.x * pitch;
      const NormState st = make_norm_state(row[to2 - 2], row[to2 - 1], row[to2 - 6], row[to2 - 5]);
      for (int r = 0; r < pitch; r += 2) norm_fwd(row[r], row[r + 1], st, flags);
    }
    __syncthreads();
    for (int q = 0; q < GG_ACC; ++q) {
      const int e = threadIdx.x + q * GG_THREADS;
      if (e >= n_out) break;
      int off, a, b;
      if (e < to2 * to2) { off = 0; a = e / to2; b = e % to2; }
      else { off = to2; a = (e - to2 * to2) / tp2; b = (e - to2 * to2) % tp2; }
      double sum = 0.0;
      for (int m = 0; m < GG_CHUNK; ++m) sum = fma((double)xs[m * pitch + off + a], (double)xs[m * pitch + off + b], sum);
      acc[q] += sum;
    }
    __syncthreads();
  }
  for (int q = 0; q < GG_ACC; ++q) {
    const int e = threadIdx.x + q * GG_THREADS;
    if (e >= n_out) break;
    if (e < to2 * to2) atomicAdd(G_obs + e, acc[q]);
    else atomicAdd(G_pred + (e - to2 * to2), acc[q]);
  }
}

// =======================================================================================
// 2. Symmetric eigen-solve of G (m <= 64) by parallel-ordered cyclic Jacobi in float64, one block.
//    Per-pair threshold |g_pq| <= eps * sqrt(g_pp g_qq) (high relative accuracy on PSD matrices;
//    exact-zero rows/columns -- the normalised last observed frame -- never rotate).
// =======================================================================================
constexpr int EIG_MAX_THREADS = 288;
constexpr double EIGF_REL = 1e-12;      // relative rotation threshold of the compile-time-size solver
constexpr int EIG_MAX_SWEEPS = 60;

// The solve is latency-bound (a 24 x 24 matrix has 288 work items per phase) and one warp spends ~3 cycles per
// instruction on it, so the item loops are spread over a few warps (NT threads: 2-3 items each); a single-warp block
// separates phases by __syncwarp.
__device__ __forceinline__ void eig_sync() {
  if (blockDim.x <= 32) __syncwarp();
  else __syncthreads();
}

// MP != 0: the padded size and the block size NT are compile-time constants, so every index computation and the item
// loops unroll completely.
template <int MP, int NT>
__device__ __forceinline__ void eig_jacobi_body(const double* __restrict__ G, int m, int k, float* __restrict__ U,
                                                float* __restrict__ S, double* __restrict__ U64,
                                                double* __restrict__ S64, int* __restrict__ info, double* sm) {
  const int mp = MP ? MP : ((m + 1) & ~1);   // even size; a padding index never rotates
  const int ld = mp + 1;         // odd pitch: column walks (stride ld doubles) are bank-conflict free
  double* A = sm;                // mp x mp, row-major with pitch ld
  double* V = A + mp * ld;       // mp x mp, same pitch
  double* cs = V + mp * ld;      // mp/2 cosines, mp/2 sines
  int* pr = reinterpret_cast<int*>(cs + mp);   // pairs p[mp/2], q[mp/2]
  int* order = pr + mp;                         // mp
  __shared__ int n_rot, step_rot[2];   // step_rot is double-buffered by step parity (reset one step ahead)
  __shared__ double floor2;            // (1e-18 * largest diagonal entry)^2: rotations below it cannot matter
  int sweeps_done = 0, total_rot = 0;
  const int tid = threadIdx.x, nthr = MP ? NT : (int)blockDim.x, half = mp / 2;

  for (int e = tid; e < mp * mp; e += nthr) {
    const int r = e / mp, c = e % mp;
    A[r * ld + c] = (r < m && c < m) ? 0.5 * (G[r * m + c] + G[c * m + r]) : 0.0;
    V[r * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  if (tid == 0) step_rot[0] = step_rot[1] = 0;
  int gstep = 0;
  eig_sync();
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(A[i * ld + i]));
    floor2 = (1e-18 * mx) * (1e-18 * mx);
  }
  eig_sync();

  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
    if (tid == 0) n_rot = 0;
    eig_sync();
    for (int step = 0; step < mp - 1; ++step, ++gstep) {
      // round-robin tournament: position 0 fixed, the others rotate
      for (int pi = tid; pi < half; pi += nthr) {
        auto player = [&](int pos) { return pos == 0 ? 0 : 1 + (pos - 1 + step) % (mp - 1); };
        int p = player(pi), q = player(mp - 1 - pi);
        if (p > q) { const int t = p; p = q; q = t; }
        const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        double c = 1.0, s = 0.0;
        // rotate iff |a_pq| > 1e-14 sqrt(a_pp a_qq)  (compared squared: no square root on the critical path)
        // and |a_pq| above an absolute floor 1e-18 max_i a_ii that keeps the solve from polishing the numerical null space
        if (q < m && apq * apq > floor2 && apq * apq > 1e-28 * fabs(app * aqq)) {
          // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = (a_qq - a_pp) / (2 a_pq), written with one sqrt and
          // one division: t = o / (dd + sign(dd) sqrt(dd^2 + o^2)), o = 2 a_pq, dd = a_qq - a_pp
          const double o = 2.0 * apq, dd = aqq - app;
          const double h = sqrt(fma(dd, dd, o * o));
          const double t = o / (dd + (dd >= 0.0 ? h : -h));
          c = rsqrt(fma(t, t, 1.0));
          s = t * c;
          atomicAdd(&step_rot[gstep & 1], 1);
        }
        cs[pi] = c; cs[half + pi] = s; pr[pi] = p; pr[half + pi] = q;
      }
      eig_sync();
      const int rotated = step_rot[gstep & 1];     // block-uniform
      if (tid == 0) { n_rot += rotated; step_rot[(gstep + 1) & 1] = 0; }
      eig_sync();
      if (rotated == 0) continue;       // nothing to do in this step (typical for the last, confirming sweep)
      // columns: A <- A J, V <- V J
#pragma unroll
      for (int e = tid; e < half * mp; e += nthr) {
        const int pi = e / mp, r = e % mp;
        const double c = cs[pi], s = cs[half + pi];
        if (s != 0.0) {
          const int p = pr[pi], q = pr[half + pi];
          const double ap = A[r * ld + p], aq = A[r * ld + q];
          A[r * ld + p] = c * ap - s * aq;
          A[r * ld + q] = s * ap + c * aq;
          const double vp = V[r * ld + p], vq = V[r * ld + q];
          V[r * ld + p] = c * vp - s * vq;
          V[r * ld + q] = s * vp + c * vq;
        }
      }
      eig_sync();
      // rows: A <- J^T A; the rotated off-diagonal pair is zero by construction and is stored as exactly zero
#pragma unroll
      for (int e = tid; e < half * mp; e += nthr) {
        const int pi = e / mp, col = e % mp;
        const double c = cs[pi], s = cs[half + pi];
        if (s != 0.0) {
          const int p = pr[pi], q = pr[half + pi];
          const double ap = A[p * ld + col], aq = A[q * ld + col];
          A[p * ld + col] = (col == q) ? 0.0 : c * ap - s * aq;
          A[q * ld + col] = (col == p) ? 0.0 : s * ap + c * aq;
        }
      }
      eig_sync();
    }
    eig_sync();
    ++sweeps_done;
    total_rot += n_rot;
    if (n_rot == 0) break;
    eig_sync();
  }
  if (info && tid == 0) { info[0] = sweeps_done; info[1] = total_rot; }

  // order eigenvalues descending (ties: lower index first) -- m <= 64, one thread
  if (tid == 0) {
    for (int i = 0; i < m; ++i) order[i] = i;
    for (int i = 0; i < m; ++i) {
      int best = i;
      for (int j = i + 1; j < m; ++j)
        if (A[order[j] * ld + order[j]] > A[order[best] * ld + order[best]]) best = j;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
  }
  eig_sync();
  for (int j = tid; j < k; j += nthr) {
    const int col = order[j];
    const double lam = A[col * ld + col];
    const double sv = sqrt(lam > 0.0 ? lam : 0.0);
    // canonical sign: the largest-magnitude component (first one on ties) is positive
    int arg = 0;
    double big = -1.0;
    for (int r = 0; r < m; ++r) {
      const double a = fabs(V[r * ld + col]);
      if (a > big) { big = a; arg = r; }
    }
    const double sign = V[arg * ld + col] < 0.0 ? -1.0 : 1.0;
    for (int r = 0; r < m; ++r) {
      const double v = sign * V[r * ld + col];
      U[r * k + j] = (float)v;
      if (U64) U64[r * k + j] = v;
    }
    S[j] = (float)sv;
    if (S64) S64[j] = sv;
  }
}

// Compile-time sizes (16 x 16 and 24 x 24, the reference's two bases): the same parallel-ordered cyclic Jacobi with TWO
// block barriers per step instead of four.  One barrier ends the rotation-parameter phase and counts the rotating
// pairs on the way (__syncthreads_count: no shared counters); then the two-sided update A <- J^T A J is applied in ONE
// phase, one thread per 2 x 2 block (row pair i, column pair j: right rotation of pair j, then left rotation of pair i,
// the same operations the column pass followed by the row pass would perform), while a second group of warps applies
// V <- V J, one thread per (VR rows, pair).  EigFast<MP>::THREADS threads: [0, (MP/2)^2) update A, [AW, AW + MP/2 * MP/VR)
// update V (304 threads for 24 x 24 with VR = 2).
// Reciprocal and reciprocal square root in float64 from the hardware seeds (rcp / rsqrt.approx.ftz.f64, ~2^-22 relative)
// and two Newton steps each: full double accuracy up to a few ulp at a third of the latency of the IEEE-rounded
// division / sqrt sequences.  The Jacobi rotation needs c^2 + s^2 = 1 to rounding, not a correctly rounded angle.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
template <int NR = 2>
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r * 0.5, fma(-x * r, r, 1.0), r);                       // ~2^-44 after the first step
  if (NR > 1) r = fma(r * 0.5, fma(-x * r, r, 1.0), r);           // rounding level
  return r;
}

template <int MP, int VR = 2>
struct EigFast {
  static constexpr int HALF = MP / 2;
  static constexpr int AW = (HALF * HALF + 31) & ~31;       // first thread of the V group (warp aligned)
  static constexpr int VROWS = MP / VR;                     // rows per pair handled by different threads (VR rows each)
  static constexpr int THREADS = AW + HALF * VROWS;
};

// PROF (diagnostic, ET_TUNE_EIG_THREADS = 3001 / 3002): thread 0 accumulates clock64() cycles per phase into info[2..13]:
// {set-up, rotation parameters, first barrier, update, second barrier, rotating steps, idle steps, idle-step cycles,
//  ordering, output}; info must then hold 14 ints.
struct EigProf {
  long long t_setup = 0, t_param = 0, t_bar1 = 0, t_upd = 0, t_bar2 = 0, t_idle = 0, t_order = 0, t_out = 0, t_begin = 0, t_wall = 0;
  int full = 0, idle = 0;
  __device__ __forceinline__ void store(int* info) const {
    info[2] = (int)t_setup; info[3] = (int)t_param; info[4] = (int)t_bar1; info[5] = (int)t_upd; info[6] = (int)t_bar2;
    info[7] = full; info[8] = idle; info[9] = (int)t_idle; info[10] = (int)t_order; info[11] = (int)t_out;
    info[12] = (int)(clock64() - t_begin);      // whole body, entry to here
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    info[13] = (int)(ns - (unsigned long long)t_wall);   // the same in nanoseconds of the global timer
  }
};

template <int MP, int NR = 2, int VR = 2, bool PROF = false>
__device__ __forceinline__ void eig_jacobi_fast(const double* __restrict__ G, int k, float* __restrict__ U,
                                                float* __restrict__ S, double* __restrict__ U64, double* __restrict__ S64,
                                                int* __restrict__ info, double* sm) {
  constexpr int m = MP, half = MP / 2, ld = MP + 1, AW = EigFast<MP, VR>::AW, VROWS = EigFast<MP, VR>::VROWS;
  double* A = sm;                // MP x MP, row-major with odd pitch
  double* V = A + MP * ld;
  double* cs = V + MP * ld;      // half cosines, half sines
  int* pr = reinterpret_cast<int*>(cs + MP);   // pairs p[half], q[half]
  int* order = pr + MP;
  __shared__ double floor2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  EigProf prof;
  long long tk = 0;
  if (PROF) {
    tk = clock64();
    prof.t_begin = tk;
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    prof.t_wall = (long long)ns;
  }
  for (int e = tid; e < MP * MP; e += nthr) {
    const int r = e / MP, c = e % MP;
    A[r * ld + c] = 0.5 * (G[r * m + c] + G[c * m + r]);
    V[r * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(A[i * ld + i]));
    floor2 = (1e-18 * mx) * (1e-18 * mx);
  }
  __syncthreads();
  if (PROF) { const long long t = clock64(); prof.t_setup = t - tk; tk = t; }
  int sweeps_done = 0, total_rot = 0;
  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
    int n_rot = 0;
    for (int step = 0; step < MP - 1; ++step) {
      bool rotating = false;
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (PROF) t0 = clock64();
      if (tid < half) {          // round-robin tournament: position 0 fixed, the others rotate
        const int pi = tid;
        auto player = [&](int pos) { return pos == 0 ? 0 : 1 + (pos - 1 + step) % (MP - 1); };
        int p = player(pi), q = player(MP - 1 - pi);
        if (p > q) { const int t = p; p = q; q = t; }
        const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
        double c = 1.0, s = 0.0;
        // rotate iff |a_pq| > EIGF_REL sqrt(a_pp a_qq) (compared squared) and above the absolute floor (see the generic
        // body).  1e-12 relative leaves the eigenvectors ~1e-12 from converged -- five orders below what the fp32 outputs
        // resolve -- and ends the iteration one to two sweeps earlier than 1e-14.
        if (apq * apq > floor2 && apq * apq > (EIGF_REL * EIGF_REL) * fabs(app * aqq)) {
          // inner rotation (|theta| <= pi/4) from the double angle: cos 2theta = |dd| / h, h = sqrt(dd^2 + o^2), hence
          // c^2 = (h + |dd|) / (2h), s = sign(dd) o / (2 h c) -- two reciprocal square roots on the dependency chain
          // (14 dependent float64 operations) instead of sqrt, division and rsqrt of the tangent form (26)
          const double o = 2.0 * apq, dd = aqq - app;
          const double rh = fast_rsqrt<NR>(fma(dd, dd, o * o));       // 1 / h
          const double c2 = fma(0.5 * fabs(dd), rh, 0.5);
          const double rc = fast_rsqrt<NR>(c2);                        // 1 / c
          c = c2 * rc;
          s = (dd >= 0.0 ? 0.5 : -0.5) * o * rh * rc;
          rotating = true;
        }
        cs[pi] = c; cs[half + pi] = s; pr[pi] = p; pr[half + pi] = q;
      }
      if (PROF) t1 = clock64();
      const int rotated = __syncthreads_count(rotating);      // barrier + number of rotating pairs, block-uniform
      if (PROF) t2 = clock64();
      n_rot += rotated;
      if (rotated == 0) {               // nothing to do in this step (typical for the last, confirming sweep)
        if (PROF) { ++prof.idle; prof.t_idle += t2 - t0; }
        continue;
      }
      if (tid < half * half) {
        const int i = tid / half, j = tid % half;
        const double ci = cs[i], si = cs[half + i], cj = cs[j], sj = cs[half + j];
        if (si != 0.0 || sj != 0.0) {
          const int pi_ = pr[i], qi = pr[half + i], pj = pr[j], qj = pr[half + j];
          const double x00 = A[pi_ * ld + pj], x01 = A[pi_ * ld + qj], x10 = A[qi * ld + pj], x11 = A[qi * ld + qj];
          // columns: X J_j
          const double t00 = cj * x00 - sj * x01, t01 = sj * x00 + cj * x01;
          const double t10 = cj * x10 - sj * x11, t11 = sj * x10 + cj * x11;
          // rows: J_i^T T; the rotated off-diagonal pair is zero by construction and is stored as exactly zero
          const bool diag = (i == j) && si != 0.0;
          A[pi_ * ld + pj] = ci * t00 - si * t10;
          A[pi_ * ld + qj] = diag ? 0.0 : ci * t01 - si * t11;
          A[qi * ld + pj] = diag ? 0.0 : si * t00 + ci * t10;
          A[qi * ld + qj] = si * t01 + ci * t11;
        }
      } else if (tid >= AW && tid < AW + half * VROWS) {
        const int e = tid - AW, pi = e / VROWS, r0 = e % VROWS;
        const double c = cs[pi], s = cs[half + pi];
        if (s != 0.0) {
          const int p = pr[pi], q = pr[half + pi];
#pragma unroll
          for (int v = 0; v < VR; ++v) {
            const int r = r0 + v * VROWS;
            const double vp = V[r * ld + p], vq = V[r * ld + q];
            V[r * ld + p] = c * vp - s * vq;
            V[r * ld + q] = s * vp + c * vq;
          }
        }
      }
      if (PROF) t3 = clock64();
      __syncthreads();
      if (PROF) {
        const long long t4 = clock64();
        ++prof.full; prof.t_param += t1 - t0; prof.t_bar1 += t2 - t1; prof.t_upd += t3 - t2; prof.t_bar2 += t4 - t3;
      }
    }
    ++sweeps_done;
    total_rot += n_rot;
    if (n_rot == 0) break;
  }
  if (info && tid == 0) { info[0] = sweeps_done; info[1] = total_rot; }
  if (PROF) tk = clock64();

  // order eigenvalues descending (ties: lower index first), canonical sign, S = sqrt(lambda)
  if (tid == 0) {
    for (int i = 0; i < m; ++i) order[i] = i;
    for (int i = 0; i < m; ++i) {
      int best = i;
      for (int j = i + 1; j < m; ++j)
        if (A[order[j] * ld + order[j]] > A[order[best] * ld + order[best]]) best = j;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
  }
  __syncthreads();
  if (PROF) { const long long t = clock64(); prof.t_order = t - tk; tk = t; }
  for (int j = tid; j < k; j += nthr) {
    const int col = order[j];
    const double lam = A[col * ld + col];
    const double sv = sqrt(lam > 0.0 ? lam : 0.0);
    int arg = 0;
    double big = -1.0;
    for (int r = 0; r < m; ++r) {
      const double a = fabs(V[r * ld + col]);
      if (a > big) { big = a; arg = r; }
    }
    const double sign = V[arg * ld + col] < 0.0 ? -1.0 : 1.0;
    for (int r = 0; r < m; ++r) {
      const double v = sign * V[r * ld + col];
      U[r * k + j] = (float)v;
      if (U64) U64[r * k + j] = v;
    }
    S[j] = (float)sv;
    if (S64) S64[j] = sv;
  }
  if (PROF) {
    __syncthreads();
    if (tid == 0 && info) { prof.t_out = clock64() - tk; prof.store(info); }
  }
}

// Second generation of the two-barrier body (same parallel-ordered cyclic Jacobi, same threshold).  Cycle counters in the
// first generation (ET_TUNE_EIG_THREADS = 3001, [B200] 24 x 24: 186 rotating steps of ~880 cycles) showed the step split
// evenly between the rotation-parameter phase of ONE warp (408 cycles, the other warps wait at the barrier) and the update
// (403 cycles: three dependent shared-memory rounds -- sines, pair indices, matrix entries -- each behind a re-derivation of
// the shared window base from a special register, and divergent skip branches), plus 26 000 cycles of single-thread
// selection sort at the end.  Here:
//  * a dedicated warp computes the rotation parameters; every other thread prepares the shared-memory ADDRESSES of its
//    work items while it waits for that warp (pair indices come from a table built once, 23 x 24 bytes, instead of the
//    tournament arithmetic), so the update is ONE round of loads, four dependent float64 levels and the stores;
//  * all shared-memory traffic of the loop uses explicit 32-bit window addresses (ld.shared / st.shared), computed once;
//  * the rotation is computed without a branch (the threshold test runs beside the chain and selects at the end) from a
//    shorter chain: 1/h = rsqrt(dd^2 + o^2) from the hardware seed and ONE Newton step (2^-44: it only sets the ANGLE),
//    c2 = 1/2 + |dd| / 2h, then (c~, s~) = (c2, +-o / 2h) * rsqrt_seed(c2) -- a vector of the right direction whose
//    length is 1 + O(2^-21) -- normalised by the series n = 1 - d/2 + 3 d^2 / 8, d = c~^2 + s~^2 - 1 (error 5 d^3 / 16 <
//    1e-18): c^2 + s^2 = 1 to rounding, 15 dependent float64 operations instead of 28;
//  * because the angle is no longer exact to the last bit, the rotated pair is stored as computed, not as an exact zero
//    (zeroing it would perturb the matrix by 2^-44 |a_pq|); convergence and accuracy are unchanged (same rotation counts
//    per sweep, residual / orthogonality at the 1e-15 level: scripts/exp/eig_gen2.py);
//  * ordering by rank counting (m threads) and sign / output by one warp per column instead of one thread.
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double raw_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// SYM: only the 2 x 2 blocks (row pair i, column pair j) with i <= j are updated and every matrix element lives at its
// canonical place [min(r, c)][max(r, c)] -- the update is bound by the throughput of the float64 pipe (a DFMA / DMUL warp
// instruction occupies it for ~11-16 cycles on this part), so half the blocks is the largest lever left; it also frees two
// warps, and the roles are laid out so that every scheduler carries the same number of float64 instructions per step.
template <int MP, int VR = 2, bool SYM = false>
struct EigFast2 {
  static constexpr int HALF = MP / 2;
  static constexpr int NA = SYM ? HALF * (HALF + 1) / 2 : HALF * HALF;   // [0, NA): 2 x 2 blocks of A
  static constexpr int AW = (NA + 31) & ~31;                // [AW, AW + HALF * VROWS): V
  static constexpr int VROWS = EigFast<MP, VR>::VROWS;
  static constexpr int PW = (AW + HALF * VROWS + 31) & ~31;  // first thread of the rotation-parameter warp
  static constexpr int THREADS = PW + 32;
};
static_assert(EigFast2<24>::AW == EigFast<24>::AW && EigFast2<16>::AW == EigFast<16>::AW, "full layout = the first generation's");

template <int MP, int VR = 2, bool PROF = false, bool SYM = false>
__device__ __forceinline__ void eig_jacobi_fast2(const double* __restrict__ G, int k, float* __restrict__ U,
                                                 float* __restrict__ S, double* __restrict__ U64, double* __restrict__ S64,
                                                 int* __restrict__ info, double* sm) {
  constexpr int m = MP, half = MP / 2, ld = MP + 1, NSTEP = MP - 1;
  constexpr int AW = EigFast2<MP, VR, SYM>::AW, VROWS = EigFast2<MP, VR, SYM>::VROWS, PW = EigFast2<MP, VR, SYM>::PW;
  constexpr int NA = EigFast2<MP, VR, SYM>::NA;
  static_assert(MP <= 32, "one lane per row in the output phase");
  double* A = sm;                // MP x MP, row-major with odd pitch
  double* V = A + MP * ld;
  double* cs = V + MP * ld;      // half cosines, half sines
  int* order = reinterpret_cast<int*>(cs + MP) + MP;   // (same layout as the first generation; its pair arrays stay unused)
  __shared__ double floor2;
  __shared__ unsigned char tab[NSTEP * MP];            // tab[step][pi] = p, tab[step][half + pi] = q  (p < q)
  const int tid = threadIdx.x, nthr = blockDim.x;
  EigProf prof;
  long long tk = 0;
  if (PROF) {
    tk = clock64();
    prof.t_begin = tk;
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    prof.t_wall = (long long)ns;
  }
  for (int e = tid; e < MP * MP; e += nthr) {
    const int r = e / MP, c = e % MP;
    A[r * ld + c] = 0.5 * (G[r * m + c] + G[c * m + r]);
    V[r * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  for (int e = tid; e < NSTEP * half; e += nthr) {     // round-robin tournament: position 0 fixed, the others rotate
    const int step = e / half, pi = e % half;
    auto player = [&](int pos) { return pos == 0 ? 0 : 1 + (pos - 1 + step) % (MP - 1); };
    int p = player(pi), q = player(MP - 1 - pi);
    if (p > q) { const int t = p; p = q; q = t; }
    tab[step * MP + pi] = (unsigned char)p;
    tab[step * MP + half + pi] = (unsigned char)q;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(A[i * ld + i]));
    floor2 = (1e-18 * mx) * (1e-18 * mx);
  }
  __syncthreads();
  // roles and loop-invariant shared-window addresses
  int role = tid < NA ? 0 : (tid >= AW && tid < AW + half * VROWS ? 1 : (tid >= PW && tid < PW + half ? 2 : 3));
  asm volatile("" : "+r"(role));      // opaque: kept in a register instead of being re-derived from %tid.x (a special-register
                                      // read on the critical path) at the top of every step
  const bool isA = role == 0, isV = role == 1, isP = role == 2;
  const bool prof_thread = PROF && tid == PW;
  int ai = isA ? tid / half : 0, aj = isA ? tid % half : 0;                            // A: 2 x 2 block (row pair ai, column pair aj)
  if (SYM && isA) {                 // item t of the upper triangle of pair indices, row by row: (0,0) .. (0,half-1), (1,1) ..
    int t = tid, i = 0;
    while (t >= half - i) { t -= half - i; ++i; }
    ai = i; aj = i + t;
  }
  const int vpi = isV ? (tid - AW) / VROWS : 0, vr0 = isV ? (tid - AW) % VROWS : 0;    // V: pair vpi, rows vr0 + v * VROWS
  const int ppi = isP ? tid - PW : 0;                                                   // parameters of pair ppi
  unsigned a_sh = (unsigned)__cvta_generic_to_shared(A), v_sh = (unsigned)__cvta_generic_to_shared(V);
  unsigned cs_sh = (unsigned)__cvta_generic_to_shared(cs), tab_sh = (unsigned)__cvta_generic_to_shared(tab);
  asm volatile("" : "+r"(a_sh), "+r"(v_sh), "+r"(cs_sh), "+r"(tab_sh));   // (opaque for the same reason: the window base is a special register)
  const unsigned adr_ci = cs_sh + 8u * (isA ? ai : vpi), adr_si = adr_ci + 8u * half;   // (V threads: their pair's c, s)
  const unsigned adr_cj = cs_sh + 8u * aj, adr_sj = adr_cj + 8u * half;
  const unsigned adr_cout = cs_sh + 8u * ppi, adr_sout = adr_cout + 8u * half;
  const double fl2 = floor2;
  // the parameter threads keep the addresses of a_pp, a_qq, a_pq of their NEXT pair in registers
  unsigned adr_pp = a_sh, adr_qq = a_sh, adr_pq = a_sh;
  auto pair_addresses = [&](unsigned p, unsigned q) {
    adr_pp = a_sh + p * (unsigned)((ld + 1) * 8);
    adr_qq = a_sh + q * (unsigned)((ld + 1) * 8);
    adr_pq = a_sh + (p * (unsigned)ld + q) * 8u;
  };
  if (isP) pair_addresses(tab[ppi], tab[half + ppi]);
  if (PROF) { const long long t = clock64(); prof.t_setup = t - tk; tk = t; }
  int sweeps_done = 0, total_rot = 0;
  for (int sweep = 0; sweep < EIG_MAX_SWEEPS; ++sweep) {
    int n_rot = 0;
    for (int step = 0; step < NSTEP; ++step) {
      const unsigned ts = tab_sh + (unsigned)(step * MP);
      bool rotating = false;
      long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (PROF) t0 = clock64();
      unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;        // addresses of this thread's work items (A: x00, x01, x10, x11; V: v_p, v_q of row vr0)
      unsigned nb0 = 0, nb1 = 0;
      if (isP) {
        const unsigned tn = tab_sh + (unsigned)((step + 1 == NSTEP ? 0 : step + 1) * MP);   // (the table wraps at the end of a sweep)
        nb0 = lds_u8(tn + ppi);
        nb1 = lds_u8(tn + half + ppi);
        const double app = lds_f64(adr_pp), aqq = lds_f64(adr_qq), apq = lds_f64(adr_pq);
        const double o = 2.0 * apq, dd = aqq - app, apq2 = apq * apq;
        const double x = fma(dd, dd, o * o);
        double rh = raw_rsqrt(x);                                   // 1 / h, h = sqrt(dd^2 + o^2)
        rh = fma(rh * 0.5, fma(-x * rh, rh, 1.0), rh);
        const double c2 = fma(0.5 * fabs(dd), rh, 0.5);             // cos^2 of the inner rotation angle (|theta| <= pi/4)
        const double kk = (dd >= 0.0 ? 0.5 : -0.5) * o * rh;
        const double rc = raw_rsqrt(c2);
        const double ct = c2 * rc, st = kk * rc;
        const double d = fma(ct, ct, fma(st, st, -1.0));
        const double nn = fma(d, fma(0.375, d, -0.5), 1.0);
        // rotate iff |a_pq| > EIGF_REL sqrt(a_pp a_qq) (compared squared) and above the absolute floor; a pair that does not
        // rotate may have produced NaN above (0 / 0) -- the selection discards it
        rotating = apq2 > fl2 && apq2 > (EIGF_REL * EIGF_REL) * fabs(app * aqq);
        sts_f64(adr_cout, rotating ? ct * nn : 1.0);
        sts_f64(adr_sout, rotating ? st * nn : 0.0);
      } else if (isA) {
        const unsigned pi_ = lds_u8(ts + ai), qi = lds_u8(ts + half + ai), pj = lds_u8(ts + aj), qj = lds_u8(ts + half + aj);
        auto at = [&](unsigned r, unsigned c) {       // SYM: the canonical place of element {r, c}
          const unsigned lo = SYM ? (r < c ? r : c) : r, hi = SYM ? (r < c ? c : r) : c;
          return a_sh + (lo * (unsigned)ld + hi) * 8u;
        };
        w0 = at(pi_, pj);
        w1 = at(pi_, qj);
        w2 = at(qi, pj);       // (SYM, diagonal block: the same place as w1 -- the block is symmetric; the later store wins)
        w3 = at(qi, qj);
      } else if (isV) {
        const unsigned vp = lds_u8(ts + vpi), vq = lds_u8(ts + half + vpi);
        w0 = v_sh + ((unsigned)(vr0 * ld) + vp) * 8u;
        w1 = v_sh + ((unsigned)(vr0 * ld) + vq) * 8u;
      }
      if (PROF) t1 = clock64();
      const int rotated = __syncthreads_count(rotating);      // barrier + number of rotating pairs, block-uniform
      if (PROF) t2 = clock64();
      if (isP) pair_addresses(nb0, nb1);
      n_rot += rotated;
      if (rotated == 0) {               // nothing to do in this step (typical for the last, confirming sweep)
        if (prof_thread) { ++prof.idle; prof.t_idle += t2 - t0; }
        continue;
      }
      if (isA) {
        const double ci = lds_f64(adr_ci), si = lds_f64(adr_si), cj = lds_f64(adr_cj), sj = lds_f64(adr_sj);
        const double x00 = lds_f64(w0), x01 = lds_f64(w1), x10 = lds_f64(w2), x11 = lds_f64(w3);
        // columns: X J_j
        const double t00 = cj * x00 - sj * x01, t01 = sj * x00 + cj * x01;
        const double t10 = cj * x10 - sj * x11, t11 = sj * x10 + cj * x11;
        // rows: J_i^T T  (an identity rotation reproduces its operands exactly; untouched blocks are not stored)
        if (si != 0.0 || sj != 0.0) {
          sts_f64(w0, ci * t00 - si * t10);
          sts_f64(w1, ci * t01 - si * t11);
          sts_f64(w2, si * t00 + ci * t10);
          sts_f64(w3, si * t01 + ci * t11);
        }
      } else if (isV) {
        const double c = lds_f64(adr_ci), s = lds_f64(adr_si);
        double vp[VR], vq[VR];
#pragma unroll
        for (int v = 0; v < VR; ++v) {
          vp[v] = lds_f64(w0 + (unsigned)(v * VROWS * ld * 8));
          vq[v] = lds_f64(w1 + (unsigned)(v * VROWS * ld * 8));
        }
        if (s != 0.0) {
#pragma unroll
          for (int v = 0; v < VR; ++v) {
            sts_f64(w0 + (unsigned)(v * VROWS * ld * 8), c * vp[v] - s * vq[v]);
            sts_f64(w1 + (unsigned)(v * VROWS * ld * 8), s * vp[v] + c * vq[v]);
          }
        }
      }
      if (PROF) t3 = clock64();
      __syncthreads();
      if (prof_thread) {
        const long long t4 = clock64();
        ++prof.full; prof.t_param += t1 - t0; prof.t_bar1 += t2 - t1; prof.t_upd += t3 - t2; prof.t_bar2 += t4 - t3;
      }
    }
    ++sweeps_done;
    total_rot += n_rot;
    if (n_rot == 0) break;
  }
  if (info && tid == 0) { info[0] = sweeps_done; info[1] = total_rot; }
  if (PROF) tk = clock64();

  // eigenvalues descending (ties: lower index first) by rank counting
  if (tid < m) order[tid] = tid;
  __syncthreads();
  if (tid < m) {
    const double li = A[tid * ld + tid];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double lj = A[j * ld + j];
      rank += (lj > li || (lj == li && j < tid)) ? 1 : 0;
    }
    order[rank] = tid;
  }
  __syncthreads();
  if (PROF) { const long long t = clock64(); prof.t_order = t - tk; tk = t; }
  // one (full) warp per output column, one lane per row: canonical sign (the largest-magnitude component, the first one on
  // ties, is positive), S = sqrt(lambda)
  const int warp = tid >> 5, lane = tid & 31, nfull = nthr >> 5;
  if (warp < nfull) {
    for (int j = warp; j < k; j += nfull) {
      const int col = order[j];
      const double v = lane < m ? V[lane * ld + col] : 0.0;
      double a = lane < m ? fabs(v) : -1.0;
      int arg = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double a2 = __shfl_xor_sync(0xffffffffu, a, o);
        const int g2 = __shfl_xor_sync(0xffffffffu, arg, o);
        if (a2 > a || (a2 == a && g2 < arg)) { a = a2; arg = g2; }
      }
      const double vbig = __shfl_sync(0xffffffffu, v, arg);
      const double w = vbig < 0.0 ? -v : v;
      if (lane < m) {
        U[lane * k + j] = (float)w;
        if (U64) U64[lane * k + j] = w;
      }
      if (lane == 0) {
        const double lam = A[col * ld + col];
        const double sv = sqrt(lam > 0.0 ? lam : 0.0);
        S[j] = (float)sv;
        if (S64) S64[j] = sv;
      }
    }
  }
  if (PROF) {
    __syncthreads();
    if (prof_thread && info) { prof.t_out = clock64() - tk; prof.store(info); }
  }
}

constexpr int EIG_DEFAULT_GEN = 3;      // generation of the two-barrier body taken by default (ET_TUNE_EIG_THREADS 2001 / 2002 force one,
                                        // the environment variable ET_EIG_GEN = 1 | 2 | 3 changes the default of the process; 3 = the second generation
                                        // with the symmetric update)
static int eig_default_gen() {
  static const int gen = [] {
    const char* e = getenv("ET_EIG_GEN");
    return (e && e[0] >= '1' && e[0] <= '3' && e[1] == 0) ? e[0] - '0' : EIG_DEFAULT_GEN;
  }();
  return gen;
}

template <int MP, int NR = 2, int VR = 2, int GEN = 1, bool PROF = false>
__global__ void __launch_bounds__(GEN >= 2 ? EigFast2<MP, VR, GEN == 3>::THREADS : EigFast<MP, VR>::THREADS) eig_jacobi_fast_kernel(const double* __restrict__ G, int k,
                                                                                   float* __restrict__ U, float* __restrict__ S,
                                                                                   double* __restrict__ U64, double* __restrict__ S64,
                                                                                   int* __restrict__ info) {
  extern __shared__ double sm[];
  if constexpr (GEN >= 2) eig_jacobi_fast2<MP, VR, PROF, GEN == 3>(G, k, U, S, U64, S64, info, sm);
  else eig_jacobi_fast<MP, NR, VR, PROF>(G, k, U, S, U64, S64, info, sm);
}

template <int MP, int NT>
__global__ void __launch_bounds__(EIG_MAX_THREADS) eig_jacobi_kernel(const double* __restrict__ G, int m, int k,
                                                                     float* __restrict__ U, float* __restrict__ S,
                                                                     double* __restrict__ U64, double* __restrict__ S64,
                                                                     int* __restrict__ info) {
  extern __shared__ double sm[];
  eig_jacobi_body<MP, NT>(G, m, k, U, S, U64, S64, info, sm);
}

// Both bases of one descriptor (16 x 16 observation and 24 x 24 prediction Gram matrices) in ONE launch: block 0 / 1
// solve them side by side on two SMs, so the pair costs what the larger solve costs.
constexpr int EIG_PAIR_THREADS = EigFast2<24>::THREADS;   // (the first generation leaves the extra warps idle)
template <int GEN>
__global__ void __launch_bounds__(EIG_PAIR_THREADS) eig_jacobi_pair_kernel(const double* __restrict__ G_a,
                                                                           const double* __restrict__ G_b, int k,
                                                                           float* __restrict__ U_a, float* __restrict__ S_a,
                                                                           float* __restrict__ U_b, float* __restrict__ S_b) {
  extern __shared__ double sm[];
  if constexpr (GEN >= 2) {
    if (blockIdx.x == 0) eig_jacobi_fast2<16, 2, false, GEN == 3>(G_a, k, U_a, S_a, nullptr, nullptr, nullptr, sm);
    else eig_jacobi_fast2<24, 2, false, GEN == 3>(G_b, k, U_b, S_b, nullptr, nullptr, nullptr, sm);
  } else {
    if (blockIdx.x == 0) eig_jacobi_fast<16>(G_a, k, U_a, S_a, nullptr, nullptr, nullptr, sm);
    else eig_jacobi_fast<24>(G_b, k, U_b, S_b, nullptr, nullptr, nullptr, sm);
  }
}

// =======================================================================================
// 3. Batched small-N SVD: one-sided (Hestenes) Jacobi on a shared-memory resident tall-skinny matrix.
//    One block per problem; X (n_b x m) is kept column-major in fp32; a warp owns a column pair per step:
//    alpha = |x_p|^2, beta = |x_q|^2, gamma = x_p . x_q by warp-shuffle reduction (fp64 partial sums),
//    then the plane rotation is applied to the two columns of X and of V (m x m).  On exit the columns of
//    V are the left singular vectors of the wide (m x n_b) view and the column norms of X the singular values.
// =======================================================================================
constexpr int SVS_THREADS = 384;
constexpr int SVS_MAX_SWEEPS = 30;

__global__ void __launch_bounds__(SVS_THREADS) svd_small_kernel(const float* __restrict__ traj,
                                                                const int64_t* __restrict__ offsets, int ld, int m,
                                                                int k, float* __restrict__ U, float* __restrict__ S) {
  extern __shared__ float smf[];
  const int mp = (m + 1) & ~1;
  float* X = smf;                       // mp columns of pitch ld
  float* V = X + (size_t)mp * ld;       // mp x mp, column-major (pitch mp)
  float* nrm = V + mp * mp;             // mp
  int* order = reinterpret_cast<int*>(nrm + mp);
  __shared__ int n_rot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = SVS_THREADS / 32;
  const int64_t r0 = offsets[blockIdx.x];
  int nb = (int)(offsets[blockIdx.x + 1] - r0);
  if (nb > ld) nb = ld;   // the host wrapper guarantees max_rows >= every n_b; never write out of bounds
  const float* src = traj + r0 * m;

  for (int e = tid; e < nb * m; e += SVS_THREADS) {
    const int r = e / m, c = e % m;
    X[(size_t)c * ld + r] = __ldg(src + e);
  }
  if (mp != m) for (int r = tid; r < nb; r += SVS_THREADS) X[(size_t)m * ld + r] = 0.f;
  for (int e = tid; e < mp * mp; e += SVS_THREADS) V[e] = (e / mp == e % mp) ? 1.f : 0.f;
  __syncthreads();

  const int half = mp / 2;
  for (int sweep = 0; sweep < SVS_MAX_SWEEPS; ++sweep) {
    if (tid == 0) n_rot = 0;
    __syncthreads();
    for (int step = 0; step < mp - 1; ++step) {
      for (int pi = warp; pi < half; pi += nwarps) {
        auto player = [&](int pos) { return pos == 0 ? 0 : 1 + (pos - 1 + step) % (mp - 1); };
        int p = player(pi), q = player(mp - 1 - pi);
        if (p > q) { const int t = p; p = q; q = t; }
        if (q >= m) continue;
        float* xp = X + (size_t)p * ld;
        float* xq = X + (size_t)q * ld;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int r = lane; r < nb; r += 32) {
          const double vp = xp[r], vq = xq[r];
          a = fma(vp, vp, a); b = fma(vq, vq, b); g = fma(vp, vq, g);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (g * g > 1e-14 * (a * b) && fabs(g) > 1e-30) {
          const double o = 2.0 * g, dd = b - a;
          const double h = sqrt(fma(dd, dd, o * o));
          const double t = o / (dd + (dd >= 0.0 ? h : -h));
          const double cd = rsqrt(fma(t, t, 1.0));
          const float c = (float)cd, s = (float)(t * cd);
          for (int r = lane; r < nb; r += 32) {
            const float vp = xp[r], vq = xq[r];
            xp[r] = c * vp - s * vq;
            xq[r] = s * vp + c * vq;
          }
          for (int r = lane; r < m; r += 32) {
            const float vp = V[p * mp + r], vq = V[q * mp + r];
            V[p * mp + r] = c * vp - s * vq;
            V[q * mp + r] = s * vp + c * vq;
          }
          if (lane == 0) atomicAdd(&n_rot, 1);
        }
      }
      __syncthreads();
    }
    if (n_rot == 0) break;
    __syncthreads();
  }

  for (int c = warp; c < m; c += nwarps) {
    double a = 0.0;
    for (int r = lane; r < nb; r += 32) { const double v = X[(size_t)c * ld + r]; a = fma(v, v, a); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) nrm[c] = (float)sqrt(a);
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < m; ++i) order[i] = i;
    for (int i = 0; i < m; ++i) {
      int best = i;
      for (int j = i + 1; j < m; ++j)
        if (nrm[order[j]] > nrm[order[best]]) best = j;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
  }
  __syncthreads();
  if (tid < k) {
    const int col = order[tid];
    int arg = 0;
    float big = -1.f;
    for (int r = 0; r < m; ++r) {
      const float a = fabsf(V[col * mp + r]);
      if (a > big) { big = a; arg = r; }
    }
    const float sign = V[col * mp + arg] < 0.f ? -1.f : 1.f;
    float* Ub = U + (size_t)blockIdx.x * m * k;
    for (int r = 0; r < m; ++r) Ub[r * k + tid] = sign * V[col * mp + r];
    S[(size_t)blockIdx.x * k + tid] = nrm[col];
  }
}

static int gram_grid() { return sm_count(); }

}  // namespace et

using namespace et;

extern "C" {

size_t et_gram_workspace_bytes(void) { return 128 + (size_t)gram_grid() * GR_REC * sizeof(double); }

int et_gram(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, int flags, double* G_obs,
            double* G_pred, void* workspace, et_stream_t stream) {
  return et_gram_init(obs, pred, n, t_obs, t_pred, flags, G_obs, G_pred, nullptr, nullptr, nullptr, nullptr, workspace, stream);
}

int et_gram_init(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, int flags, double* G_obs,
                 double* G_pred, float* pred_norm, float* ori, float* rot, float* sca, void* workspace,
                 et_stream_t stream) {
  ET_REQUIRE(!pred_norm || pred, ET_ERR_BADARG, "et_gram_init: pred_norm requested without pred");
  ET_REQUIRE(!pred_norm || aligned16(pred_norm), ET_ERR_ALIGN, "et_gram_init: pred_norm must be 16-byte aligned");
  ET_REQUIRE(!rot || aligned16(rot), ET_ERR_ALIGN, "et_gram_init: rot must be 16-byte aligned");
  const bool extras = pred_norm || ori || rot || sca;
  ET_REQUIRE(n >= 0, ET_ERR_BADARG, "et_gram: n < 0");
  ET_REQUIRE(t_obs >= 1 && t_obs <= ET_MAX_T && (!pred || (t_pred >= 1 && t_pred <= ET_MAX_T)), ET_ERR_UNSUPPORTED,
             "et_gram: T outside [1, %d]", ET_MAX_T);
  ET_REQUIRE(!flags || t_obs >= 3, ET_ERR_UNSUPPORTED, "et_gram: normalisation needs T_obs >= 3");
  if (n == 0) return ET_OK;
  ET_REQUIRE((obs && G_obs) || n == 0, ET_ERR_BADARG, "et_gram: obs / G_obs null");
  ET_REQUIRE(!pred || G_pred, ET_ERR_BADARG, "et_gram: pred given but G_pred null");
  ET_REQUIRE(aligned16(obs) && aligned16(pred), ET_ERR_ALIGN, "et_gram: trajectory pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (t_obs == 8 && (!pred || t_pred == 12)) {
    ET_REQUIRE(workspace, ET_ERR_BADARG, "et_gram: workspace of et_gram_workspace_bytes() zero-initialised bytes required");
    ET_REQUIRE(aligned16(workspace), ET_ERR_ALIGN, "et_gram: workspace must be 16-byte aligned");
    int grid = gram_grid();
    const int64_t need = ((n + 31) / 32 + GR_WARPS - 1) / GR_WARPS;
    if (grid > need) grid = (int)need;
    unsigned* ctr = reinterpret_cast<unsigned*>(workspace);
    double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 128);
    cudaError_t ce;
    auto launch = [&](auto kern) {      // the attribute is per device: set it on every call (cheap)
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GR_SMEM);
      if (e != cudaSuccess) return e;
      return launch_cooperative(kern, dim3(grid), dim3(GR_WARPS * 32), GR_SMEM, st, obs, pred, n, flags, G_obs, G_pred, ctr, parts,
                                pred_norm, ori, rot, sca);
    };
    switch (tune_get(ET_TUNE_GRAM_UNROLL)) {
      case 1: ce = launch(gram_fast<1>); break;
      case 4: ce = launch(gram_fast<4>); break;
      case 8: ce = launch(gram_fast<8>); break;
      default: ce = launch(gram_fast<2>); break;
    }
    if (ce != cudaSuccess) return fail(ET_ERR_CUDA, "gram_fast: cooperative launch: %s", cudaGetErrorString(ce));
    return check_launch("gram_fast");
  }
  if (extras) {     // other shapes: the state and the normalised futures come from the stand-alone kernels
    int rc = et_norm_params(obs, n, t_obs, flags, ori, rot, sca, stream);
    if (rc) return rc;
    if (pred_norm && (rc = et_normalize(pred, n, t_pred, flags, ori, rot, sca, pred_norm, stream))) return rc;
  }
  const int to2 = 2 * t_obs, tp2 = pred ? 2 * t_pred : 0;
  const size_t smem = (size_t)GG_CHUNK * (to2 + tp2) * sizeof(float);
  ET_REQUIRE(to2 * to2 + tp2 * tp2 <= GG_THREADS * GG_ACC, ET_ERR_UNSUPPORTED, "et_gram: shape too large");
  int64_t grid = (n + GG_CHUNK - 1) / GG_CHUNK;
  if (grid > 2 * sm_count()) grid = 2 * sm_count();
  gram_generic<<<(unsigned)grid, GG_THREADS, smem, st>>>(obs, pred, n, to2, tp2, flags, G_obs, G_pred);
  return check_launch("gram_generic");
}

int et_eig_jacobi(const double* G, int m, int k, float* U, float* S, double* U64, double* S64, int* info,
                  et_stream_t stream) {
  ET_REQUIRE(G && U && S, ET_ERR_BADARG, "et_eig_jacobi: null pointer");
  ET_REQUIRE(m >= 1 && m <= 2 * ET_MAX_T, ET_ERR_UNSUPPORTED, "et_eig_jacobi: m = %d outside [1, %d]", m, 2 * ET_MAX_T);
  ET_REQUIRE(k >= 1 && k <= m, ET_ERR_BADARG, "et_eig_jacobi: k = %d outside [1, m = %d]", k, m);
  const int mp = (m + 1) & ~1;
  const size_t smem = (size_t)(2 * mp * (mp + 1) + mp) * sizeof(double) + (size_t)2 * mp * sizeof(int);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(eig_jacobi_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(ET_ERR_CUDA, "eig_jacobi_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  cudaStream_t st = as_stream(stream);
  // one work item per thread (measured on B200, 24 x 24: 403 / 211 / 184 / 158 us with 32 / 96 / 144 / 288 threads;
  // 16 x 16: 117 / 90 / 74 us with 32 / 64 / 128; results bit-identical)
  // 16 x 16 / 24 x 24 take the two-barrier body (eig_jacobi_fast: 24 x 24 in ~95 us); ET_TUNE_EIG_THREADS selects the
  // four-barrier body with that many threads for A/B runs (32 = the single-warp variant)
  int nt = tune_get(ET_TUNE_EIG_THREADS);
  if (nt == 0) nt = 2000 + eig_default_gen();
  // 2001 / 2002: first / second generation of the two-barrier body; 3001 / 3002: the same with phase cycle counters in
  // info[2..13] (diagnostic: info must hold 14 ints)
  ET_REQUIRE((nt != 3001 && nt != 3002 && nt != 3003) || (info && (m == 16 || m == 24)), ET_ERR_BADARG, "et_eig_jacobi: profiling variants need info[14] and m = 16 / 24");
  if (m == 16 && nt == 2001) {
    eig_jacobi_fast_kernel<16><<<1, EigFast<16>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 2001) {
    eig_jacobi_fast_kernel<24><<<1, EigFast<24>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 16 && nt == 2002) {
    eig_jacobi_fast_kernel<16, 2, 2, 2><<<1, EigFast2<16>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 2002) {
    eig_jacobi_fast_kernel<24, 2, 2, 2><<<1, EigFast2<24>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 16 && nt == 2003) {
    eig_jacobi_fast_kernel<16, 2, 2, 3><<<1, EigFast2<16, 2, true>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 2003) {
    eig_jacobi_fast_kernel<24, 2, 2, 3><<<1, EigFast2<24, 2, true>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 3003) {
    eig_jacobi_fast_kernel<24, 2, 2, 3, true><<<1, EigFast2<24, 2, true>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 16 && nt == 3001) {
    eig_jacobi_fast_kernel<16, 2, 2, 1, true><<<1, EigFast<16>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 3001) {
    eig_jacobi_fast_kernel<24, 2, 2, 1, true><<<1, EigFast<24>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 16 && nt == 3002) {
    eig_jacobi_fast_kernel<16, 2, 2, 2, true><<<1, EigFast2<16>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 3002) {
    eig_jacobi_fast_kernel<24, 2, 2, 2, true><<<1, EigFast2<24>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 1001) {      // A/B variants of the two-barrier body (measured on B200, default = 2 Newton steps
    // per rsqrt, 2 rows per V thread: 134 us): one row per V thread 140 us (same bits), four rows 138 us (same bits), one
    // Newton step 129 us but eigenvectors only ~1e-9 from the two-step result -- not taken
    eig_jacobi_fast_kernel<24, 2, 1><<<1, EigFast<24, 1>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (m == 24 && nt == 1003) {
    eig_jacobi_fast_kernel<24, 1, 2><<<1, EigFast<24, 2>::THREADS, smem, st>>>(G, k, U, S, U64, S64, info);
  } else if (mp == 16) {
    if (nt == 32) eig_jacobi_kernel<16, 32><<<1, 32, smem, st>>>(G, m, k, U, S, U64, S64, info);
    else eig_jacobi_kernel<16, 128><<<1, 128, smem, st>>>(G, m, k, U, S, U64, S64, info);
  } else if (mp == 24) {
    if (nt == 32) eig_jacobi_kernel<24, 32><<<1, 32, smem, st>>>(G, m, k, U, S, U64, S64, info);
    else if (nt == 96) eig_jacobi_kernel<24, 96><<<1, 96, smem, st>>>(G, m, k, U, S, U64, S64, info);
    else eig_jacobi_kernel<24, 288><<<1, 288, smem, st>>>(G, m, k, U, S, U64, S64, info);
  } else {
    int threads = ((mp / 2) * mp / 9 + 31) / 32 * 32;     // ~9 work items per thread and phase
    if (threads < 32) threads = 32;
    if (threads > EIG_MAX_THREADS) threads = EIG_MAX_THREADS;
    eig_jacobi_kernel<0, 0><<<1, threads, smem, st>>>(G, m, k, U, S, U64, S64, info);
  }
  return check_launch("eig_jacobi_kernel");
}

int et_eig_jacobi_pair(const double* G_a, int m_a, const double* G_b, int m_b, int k, float* U_a, float* S_a, float* U_b,
                       float* S_b, et_stream_t stream) {
  ET_REQUIRE(G_a && G_b && U_a && S_a && U_b && S_b, ET_ERR_BADARG, "et_eig_jacobi_pair: null pointer");
  ET_REQUIRE(k >= 1 && k <= m_a && k <= m_b, ET_ERR_BADARG, "et_eig_jacobi_pair: k = %d outside [1, min(m)]", k);
  if (m_a == 16 && m_b == 24) {
    const size_t smem = (size_t)(2 * 24 * 25 + 24) * sizeof(double) + (size_t)2 * 24 * sizeof(int);
    int gen = tune_get(ET_TUNE_EIG_THREADS);
    gen = gen >= 2001 && gen <= 2003 ? gen - 2000 : eig_default_gen();
    if (gen == 3) eig_jacobi_pair_kernel<3><<<2, EigFast2<24, 2, true>::THREADS, smem, as_stream(stream)>>>(G_a, G_b, k, U_a, S_a, U_b, S_b);
    else if (gen == 2) eig_jacobi_pair_kernel<2><<<2, EIG_PAIR_THREADS, smem, as_stream(stream)>>>(G_a, G_b, k, U_a, S_a, U_b, S_b);
    else eig_jacobi_pair_kernel<1><<<2, EIG_PAIR_THREADS, smem, as_stream(stream)>>>(G_a, G_b, k, U_a, S_a, U_b, S_b);
    return check_launch("eig_jacobi_pair_kernel");
  }
  int rc = et_eig_jacobi(G_a, m_a, k, U_a, S_a, nullptr, nullptr, nullptr, stream);   // other shapes: two launches
  if (rc) return rc;
  return et_eig_jacobi(G_b, m_b, k, U_b, S_b, nullptr, nullptr, nullptr, stream);
}

int et_svd_small(const float* traj, const int64_t* offsets, int batch, int64_t max_rows, int t, int k, float* U,
                 float* S, et_stream_t stream) {
  ET_REQUIRE(batch >= 0 && max_rows >= 0, ET_ERR_BADARG, "et_svd_small: bad batch / max_rows");
  ET_REQUIRE(t >= 1 && t <= ET_MAX_T, ET_ERR_UNSUPPORTED, "et_svd_small: T = %d outside [1, %d]", t, ET_MAX_T);
  const int m = 2 * t;
  ET_REQUIRE(k >= 1 && k <= m, ET_ERR_BADARG, "et_svd_small: k = %d outside [1, %d]", k, m);
  if (batch == 0) return ET_OK;
  ET_REQUIRE(traj && offsets && U && S, ET_ERR_BADARG, "et_svd_small: null pointer");
  const int mp = (m + 1) & ~1;
  const int ld = (int)max_rows | 1;   // odd pitch
  const size_t smem = ((size_t)mp * ld + (size_t)mp * mp + mp) * sizeof(float) + (size_t)mp * sizeof(int);
  ET_REQUIRE(smem <= 200 * 1024, ET_ERR_UNSUPPORTED,
             "et_svd_small: %lld rows x %d columns do not fit shared memory (use et_gram + et_eig_jacobi)",
             (long long)max_rows, m);
  cudaError_t e = cudaFuncSetAttribute(svd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "svd_small_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  svd_small_kernel<<<batch, SVS_THREADS, smem, as_stream(stream)>>>(traj, offsets, ld, m, k, U, S);
  return check_launch("svd_small_kernel");
}

}  // extern "C"
