// k-means over eigen-coefficients (reference: EigenTrajectory/kmeans.py, class BatchKMeans).
//
// Layout: data (l, d, N) and centroids (l, d, K), both contiguous fp32 exactly as the reference holds them,
// so a warp reads d coalesced 128-byte lines per 32 points: 4 d algorithmic bytes per point and iteration.
//
// The assignment reproduces the reference's CPU arithmetic bit for bit (kmeans.py:59-76,143-158):
//   dot  = one ascending fp32 FMA chain from 0                       (what MKL sgemm does for d <= 16)
//   sim  = fl(fl(fl(2 dot) - |a|^2) - |b_j|^2)                       (in-place mul_, sub_, sub_)
//   |v|^2 = separately rounded squares summed in ATen's order for a reduction over dim -2:
//           columns inside a full 32-wide block sequentially, tail columns with four interleaved partial
//           sums ((acc0 + acc1) + acc2) + acc3 where acc0 also takes the remainder rows; a tensor of 4..7
//           columns sums its first four columns sequentially            (torch 2.11 CPU, measured)
//   label = arg-max, lowest index on ties, NaN wins (torch.max).
// The similarity scan runs on packed fp32 pairs (FFMA2 / FADD2: two centroids per instruction, every half an ordinary
// IEEE operation).  The update half accumulates without atomics in lane-private fp32 records (<= 64 points each) that
// are folded per warp in fp32 row sums and above the warp in float64, always in the same order: centroid sums are
// run-to-run reproducible and far more accurate than the reference's fp32 sum over all N.
// et_kmeans_lloyd runs the whole Lloyd loop in one persistent cooperative launch; et_kmeans_lloyd_sharded adds the
// multi-GPU exchange over peer memory inside the same kernel.
#include "et_common.cuh"

namespace et {

__device__ __forceinline__ bool col_is_sequential(int64_t idx, int64_t ncols) {
  if (ncols >= 4 && ncols < 8) return idx < 4;
  return idx < 32 * (ncols / 32);
}

// sum of squares of v[0..d) in the order torch's CPU sum(dim=-2) uses for this output column
template <int DMAX>
__device__ __forceinline__ float sumsq_torch_order(const float (&v)[DMAX], int d, bool sequential) {
  float sq[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; ++i) sq[i] = __fmul_rn(v[i], v[i]);
  if (sequential) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < DMAX; ++i)
      if (i < d) acc = __fadd_rn(acc, sq[i]);
    return acc;
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int nfull = d >> 2;
#pragma unroll
  for (int g = 0; g < DMAX / 4; ++g) {
    if (g < nfull) {
      a0 = __fadd_rn(a0, sq[4 * g]);
      a1 = __fadd_rn(a1, sq[4 * g + 1]);
      a2 = __fadd_rn(a2, sq[4 * g + 2]);
      a3 = __fadd_rn(a3, sq[4 * g + 3]);
    }
  }
#pragma unroll
  for (int i = 0; i < DMAX; ++i)
    if (i >= 4 * nfull && i < d) a0 = __fadd_rn(a0, sq[i]);
  return __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
}

// best similarity / label of one point against the staged centroids (cs: [d][kpitch], bn: [kpad]).
// Four centroids are evaluated at a time (four independent FMA chains, one broadcast LDS.128 per data row); each
// similarity is still the reference's exact op sequence and the arg-max scan stays in ascending j.  Columns
// ncols..kpad-1 are padding (centroid 0, |b|^2 = +inf => similarity -inf, never selected).
// EXACT: d == DMAX is known at compile time (no per-row guards).  NANAWARE: torch.max semantics when a similarity
// can be NaN (a NaN centroid from an empty cluster, or non-finite data); the common finite case skips those tests.
template <int DMAX, bool EXACT, bool NANAWARE>
__device__ __forceinline__ void best_centroid(const float (&a)[DMAX], int d, float anorm, const float* cs, int kpitch,
                                              const float* bn, int ncols, int kpad, float& best, int& label) {
  best = NANAWARE ? 0.f : -INFINITY;
  label = 0;
  bool first = true;
  for (int j0 = 0; j0 < kpad; j0 += 4) {
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      if (EXACT || i < d) {
        const float4 c4 = *reinterpret_cast<const float4*>(cs + i * kpitch + j0);
        dot[0] = fmaf(a[i], c4.x, dot[0]);
        dot[1] = fmaf(a[i], c4.y, dot[1]);
        dot[2] = fmaf(a[i], c4.z, dot[2]);
        dot[3] = fmaf(a[i], c4.w, dot[3]);
      }
    }
    const float4 b4 = *reinterpret_cast<const float4*>(bn + j0);
    const float bnv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float y = __fsub_rn(__fsub_rn(__fmul_rn(dot[u], 2.0f), anorm), bnv[u]);
      if (NANAWARE) {
        if (j0 + u < ncols && (first || y > best || (y != y && best == best))) { best = y; label = j0 + u; }
        first = false;
      } else {
        if (y > best) { best = y; label = j0 + u; }
      }
    }
  }
}

// Scalar scan against NEGATED centroid norms (nbn), best similarity only: the seeding kernel's path for a point whose own
// norm is outside the packed fast path's range.
template <int DMAX, bool EXACT>
__device__ __forceinline__ void best_centroid_neg(const float (&a)[DMAX], int d, float anorm, const float* cs, int kpitch,
                                                  const float* nbn, int kpad, float& best) {
  best = -INFINITY;
  for (int j = 0; j < kpad; ++j) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < DMAX; ++i)
      if (EXACT || i < d) dot = fmaf(a[i], cs[i * kpitch + j], dot);
    const float y = __fadd_rn(__fsub_rn(__fmul_rn(dot, 2.0f), anorm), nbn[j]);
    if (y > best) best = y;
  }
}

// Fast path for PPL points of one lane at once (finite data and centroids only; the caller checks).  Two centroids
// per instruction: the data value is duplicated into an fp32 pair and multiplied with the pair (c_j, c_j+1) by one
// FFMA2, so the 6 x K multiply-adds of a point cost 3 K instructions; 2 dot, - |a|^2, - |b|^2 are FMUL2 / FADD2.  Each
// half of a packed operation is an ordinary IEEE fp32 operation, hence every similarity keeps the reference's exact
// rounding sequence (x - y == x + (-y) bit for bit; the centroid norms are staged negated in nbn).
// KPAD > 0: the padded cluster count is known at compile time and the scan is fully unrolled (shared-memory operands
// become immediate offsets); KPAD == 0: run-time kpad.  LABELS = false keeps only the best similarity (seeding).
template <int DMAX, bool EXACT, int PPL, int KPAD, bool LABELS, bool TOURN = false>
__device__ __forceinline__ void best_centroid_packed(const float (&a)[PPL][DMAX], int d, const float (&anorm)[PPL],
                                                     const float* cs, int kpitch, const float* nbn, int kpad_rt,
                                                     float (&best)[PPL], int (&label)[PPL]) {
  f32x2_t ad[PPL][DMAX], nan2[PPL];
#pragma unroll
  for (int u = 0; u < PPL; ++u) {
#pragma unroll
    for (int i = 0; i < DMAX; ++i) ad[u][i] = pack2(a[u][i], a[u][i]);
    nan2[u] = pack2(-anorm[u], -anorm[u]);
    best[u] = -INFINITY;
    label[u] = 0;
  }
  const f32x2_t two = pack2(2.0f, 2.0f), zero = pack2(0.f, 0.f);
  const int kpad = KPAD > 0 ? KPAD : kpad_rt;
#pragma unroll
  for (int j0 = 0; j0 < kpad; j0 += 4) {
    f32x2_t dot[PPL][2];
#pragma unroll
    for (int u = 0; u < PPL; ++u) dot[u][0] = dot[u][1] = zero;
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
      if (EXACT || i < d) {
        const ulonglong2 c4 = *reinterpret_cast<const ulonglong2*>(cs + i * kpitch + j0);
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          dot[u][0] = fma2(ad[u][i], c4.x, dot[u][0]);
          dot[u][1] = fma2(ad[u][i], c4.y, dot[u][1]);
        }
      }
    }
    const ulonglong2 nb4 = *reinterpret_cast<const ulonglong2*>(nbn + j0);
#pragma unroll
    for (int u = 0; u < PPL; ++u) {
      float y0, y1, y2, y3;
      unpack2(add2(add2(mul2(dot[u][0], two), nan2[u]), nb4.x), y0, y1);
      unpack2(add2(add2(mul2(dot[u][1], two), nan2[u]), nb4.y), y2, y3);
      if (LABELS && !TOURN) {             // plain ascending scan
        if (y0 > best[u]) { best[u] = y0; label[u] = j0; }
        if (y1 > best[u]) { best[u] = y1; label[u] = j0 + 1; }
        if (y2 > best[u]) { best[u] = y2; label[u] = j0 + 2; }
        if (y3 > best[u]) { best[u] = y3; label[u] = j0 + 3; }
      } else if (LABELS) {
        // (measured on B200: 21.75 us per iteration against 21.5 us for the plain scan above -- the kernel is bound by
        // issue slots and the FMA pipe, not by this dependency chain; kept as an A/B switch, ET_TUNE_KM_TOURNAMENT)
        // arg-max of the group as a two-level tournament (a later column wins only if strictly greater, so ties keep
        // the lowest index exactly like the ascending scan), then ONE dependent compare/select against the running
        // best: the serial chain through `best` is K/4 steps long instead of K
        const bool p01 = y1 > y0, p23 = y3 > y2;
        const float v01 = p01 ? y1 : y0, v23 = p23 ? y3 : y2;
        const int i01 = p01 ? j0 + 1 : j0, i23 = p23 ? j0 + 3 : j0 + 2;
        const bool pg = v23 > v01;
        const float vg = pg ? v23 : v01;
        const int ig = pg ? i23 : i01;
        if (vg > best[u]) { best[u] = vg; label[u] = ig; }
      } else {
        best[u] = fmaxf(best[u], fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)));
      }
    }
  }
}

constexpr int KM_WARPS = 4;                 // warps per block of the seeding kernel and the default assign kernel
constexpr int KM_THREADS = KM_WARPS * 32;
constexpr float KM_FAST_NORM_MAX = 1.0e37f;   // squared norms up to here take the packed fast path
constexpr int KM_FOLD_GROUP = 4;     // blocks that share the grid fold of the whole-fit kernel (each folds a quarter of the entries)
constexpr int KM_FLUSH_EVERY = 64;   // batches of 32 points a lane accumulates in fp32 before folding into fp64

// Whole-fit mode of the assign kernel: with cent_out != null the kernel runs the complete Lloyd loop of
// BatchKMeans.fit (kmeans.py:226-239) -- assign + accumulate, grid fold, centroid division, error test -- for up to
// max_iter iterations without returning to the host, then one more scan that writes the labels of the last assignment.
struct KmLloyd {
  int max_iter;
  double tol;
  float* cent_out;          // (l,d,K) centroids after the last update
  double* totals;           // scratch: l * (K (d+1) + 1) folded sums, then l per-entry errors
  size_t part_stride;       // doubles between the two alternating partial arrays of the whole-fit mode
  double* gtot;             // scratch: one folded record per group of KM_FOLD_GROUP blocks
  unsigned* gctr;           // one arrival counter per group (zeroed before the launch)
  double* err;              // [1] sum (old - new)^2 of the last iteration (over all l, as calculate_error)
  int32_t* status;          // [2] {converged, iterations done}
  double* simsum_last;      // (l) sum of best similarities of the last assignment
  int64_t* labels_final;    // (l,N) optional
  // Row-sharded fit over peer memory (world > 1): every rank runs this kernel on its rows; after the local grid fold
  // the ranks exchange their folded records by direct stores into each other's exchange buffers (NVLink / NVSwitch
  // peer mappings) and a per-iteration flag, then every rank adds the `world` records in rank order -- the all-reduce
  // of the reference-free sharding scheme, done inside the kernel, no NCCL call and no host in the loop.
  int rank, world;
  unsigned char* const* xchg;   // DEVICE array [world]: base of every rank's exchange buffer as mapped on THIS rank
  unsigned stamp_base;          // flags of this call are stamp_base (ready) and stamp_base + 1 + iteration
  // Row shards of one data set: the |a|^2 summation order of a column follows its GLOBAL index and the GLOBAL column
  // count (as one unsharded tensor would be summed), so sharded and unsharded labels are the same bits.
  int64_t col_offset;           // global index of this shard's column 0
  int64_t n_global;             // columns over all shards (0: this launch holds all of them)
};

// Exchange buffer of one rank: uint32 ready[world] at byte 0 (the "this rank has entered the call" stamps), then at byte
// 256 slot[2][world][l * (K (d+1) + 1)] 64-bit payload words, each stored as TWO 8-byte packets {32 payload bits, 32-bit
// stamp} (two parities: iteration i + 1 writes the other half while a slow peer may still read iteration i).  A packet
// is one naturally atomic 8-byte store, so the stamp travels WITH the data: the receiver polls the packet itself and no
// fence or separate flag (one more NVLink round trip each) sits between "record stored" and "record usable".
constexpr int KM_XCHG_HEADER = 256;
constexpr int KM_XCHG_MAX_WORLD = 16;
__device__ __forceinline__ unsigned* xchg_ready(unsigned char* base) { return reinterpret_cast<unsigned*>(base); }
// first packet of payload word `w` of rank r's record (slot_words payload words per rank and parity)
__device__ __forceinline__ unsigned long long* xchg_packets(unsigned char* base, int parity, int world, int r, size_t slot_words,
                                                           size_t w) {
  return reinterpret_cast<unsigned long long*>(base + KM_XCHG_HEADER) + (((size_t)parity * world + r) * slot_words + w) * 2;
}
__device__ __forceinline__ void ll_store(unsigned long long* pk, unsigned long long payload, unsigned stamp) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(pk), "r"((unsigned)payload), "r"(stamp) : "memory");
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(pk + 1), "r"((unsigned)(payload >> 32)), "r"(stamp) : "memory");
}
// Spin until both packets carry `stamp`; a peer that never arrives traps the kernel after ~10 s instead of hanging it.
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* pk, unsigned stamp) {
  unsigned lo, hi, s0, s1, spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(s0) : "l"(pk) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(hi), "=r"(s1) : "l"(pk + 1) : "memory");
    if (s0 == stamp && s1 == stamp) break;
    if ((++spins & 0x3ffu) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t0 == 0) t0 = t1;
      else if (t1 - t0 > 10000000000ull) __trap();
    }
  }
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// One flag per peer behind ONE system-scope fence: the fence orders everything before it (this block's record stores,
// made visible to this thread by the preceding barrier) ahead of all the relaxed flag stores that follow.  A release
// store per peer would drain the outstanding NVLink writes once per peer -- measured ~2 us each, 16 us per Lloyd
// iteration at 8 GPUs.
__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Spin until all `world` stamps at flags[0..world) have reached `stamp` (wrap-safe); a peer that never arrives traps
// the kernel after ~10 s instead of hanging the device.
__device__ __forceinline__ void xchg_wait_all(const unsigned* flags, int world, unsigned stamp) {
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int r = 0; r < world; ++r) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(flags + r) - stamp) < 0) {
      if ((++spins & 0x3ffu) == 0) {          // look at the clock every 1024 polls only
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) __trap();
      }
    }
  }
}

// Reusable grid barrier for the whole-fit mode: ctr[0] counts arrivals monotonically (barrier number `phase`, 1-based,
// completes at phase * nblocks).  km_barrier_exit() restores the zero state once every block has left its last spin.
__device__ __forceinline__ void km_barrier(unsigned* ctr, unsigned nblocks, unsigned phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release-arrive / acquire-poll at gpu scope (cumulative over the block's writes ordered by the bar.sync above)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned target = phase * nblocks;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
__device__ __forceinline__ void km_barrier_exit(unsigned* ctr, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&ctr[1], 1u) == nblocks - 1) {
    ctr[0] = 0u;
    ctr[1] = 0u;
    __threadfence();
  }
}

// workspace: four uint32 barrier counters (zero on entry / exit) padded to 128 B, then gridDim.x * gridDim.y partial
// records of (d*K + K + 1) doubles (then the whole-fit scratch).  Cooperative launch.
//
// Centroid accumulation without atomics: every lane owns a private fp32 record [cluster][d sums, count] in shared
// memory.  Coordinates are kept as fp32 pairs laid out [cluster][pair][lane] (a warp's 64-bit read-modify-write covers
// 256 contiguous bytes: conflict-free, and one FADD2 updates two coordinates), the odd coordinate (if any) and the
// count as singles [cluster][single][lane].  A lane adds at most KM_FLUSH_EVERY points into its record before the warp
// folds its 32 lane records (fp32 row sums, then fp64: see flush below) in a fixed order, so totals are reproducible
// and fp32 never carries more than 32 * KM_FLUSH_EVERY points.  KPAD: compile-time padded cluster count of the packed
// scan (0 = run time).
// SHARE: lanes per record slot.  1: every lane owns a private record (LANES = 32 columns per row).  2: lanes i and i + 16
// share a column and update it in two half-warp phases -- half the shared memory per warp, which is what lets 16 warps
// (instead of 12) live on an SM with the reference shape's 140-float records.
template <int DMAX, int KMAX, int WARPS, bool EXACT, int KPAD, bool TOURN = false, int SHARE = 1>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 1) kmeans_assign_kernel(
    const float* __restrict__ data, const float* __restrict__ centroids, int d, int64_t n, int k,
    int64_t* __restrict__ labels, float* __restrict__ maxsims, double* __restrict__ sums, double* __restrict__ counts,
    double* __restrict__ simsum, unsigned* __restrict__ barrier_ctr, double* __restrict__ partials,
    const int32_t* __restrict__ status, const int64_t* __restrict__ labels_in, const KmLloyd fit) {
  if (status && status[0] != 0) return;   // converged on an earlier iteration: nothing to do (uniform over the grid)
  constexpr int RECMAX = KMAX * (DMAX + 1);
  constexpr int NJP = (KMAX * (DMAX / 2) + 31) / 32;        // record rows per lane: coordinate pairs ...
  constexpr int NJS = (KMAX * 2 + 31) / 32;                 // ... and singles (odd coordinate, count)
  constexpr int THREADS = WARPS * 32;
  constexpr int LANES = 32 / SHARE;                         // record columns per row
  extern __shared__ __align__(16) unsigned char km_smem[];
  float* cs = reinterpret_cast<float*>(km_smem);            // [DMAX][KMAX]
  float* bn = cs + DMAX * KMAX;                             // [KMAX]   |b_j|^2
  float* nbn = bn + KMAX;                                   // [KMAX]  -|b_j|^2 (packed scan)
  float* csn = nbn + KMAX;                                  // [DMAX][KMAX] updated centroids (whole-fit mode)
  double* blk = reinterpret_cast<double*>(csn + DMAX * KMAX);   // [rec + 1]
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (EXACT) d = DMAX;                                      // lets the record layout below fold to constants
  const int rec = k * (d + 1);
  // (blk is padded to an even number of doubles: the records below are read with 128-bit accesses)
  float* lanerec = reinterpret_cast<float*>(blk + ((RECMAX + 2) & ~1)) + (size_t)warp * rec * LANES;   // rec * LANES floats
  const bool whole_fit = fit.cent_out != nullptr;
  const bool accumulate = sums != nullptr || whole_fit;
  if (whole_fit && fit.world > 1 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    // tell every rank that this rank has entered the call (so it has finished reading the previous call's slots)
    __threadfence_system();
    for (int p = 0; p < fit.world; ++p) st_relaxed_sys(xchg_ready(fit.xchg[p]) + fit.rank, fit.stamp_base);
  }
  const int npair = d >> 1, nsingle = (d & 1) + 1;          // per cluster: d/2 coordinate pairs, then (odd coordinate,) count
  float* lanesingle = lanerec + (size_t)k * npair * 2 * LANES;
  __shared__ int nan_centroid;
  __shared__ double red_s[WARPS];

  for (int e = tid; e < 2 * DMAX * KMAX + 2 * KMAX; e += THREADS) cs[e] = 0.f;   // cs, bn, nbn, csn (contiguous), padding included
  __syncthreads();
  if (centroids)
    for (int e = tid; e < d * k; e += THREADS) cs[(e / k) * KMAX + (e % k)] = __ldg(centroids + (int64_t)l * d * k + e);
  if (accumulate)
    for (int e = lane; e < rec * LANES; e += 32) lanerec[e] = 0.f;
  const int kpad = (k + 3) & ~3;
  const float* dl = data + (int64_t)l * d * n;
  const int out_rec = rec + 1;
  const unsigned nblocks = gridDim.x * gridDim.y;
  unsigned phase = 0, gphase = 1;
  double accp[NJP][2], accs[NJS];
  double sim_acc = 0.0;

  // Fold the 32 lane records of this warp: lane i owns record rows i, i+32, ... (a row = the 32 lane columns of one
  // coordinate pair, or of one single); it sums the columns of a row in fp32 with 128-bit loads and FADD2 -- one
  // instruction per element instead of a conversion and an fp64 addition each -- starting at a lane-dependent column
  // (conflict-free, and a fixed order per row), and only the row sums go to the fp64 registers.  fp32 therefore
  // carries at most 32 * KM_FLUSH_EVERY points; everything above that (warps, blocks, grid, iterations) is fp64.
  float sim32 = 0.f;            // best similarities of the points since the last flush (<= KM_FLUSH_EVERY per lane)
  auto flush = [&]() {
    sim_acc += (double)sim32;
    sim32 = 0.f;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NJP; ++j) {
      const int rho = lane + 32 * j;
      if (rho < k * npair) {
        ulonglong2* row = reinterpret_cast<ulonglong2*>(lanerec + rho * 2 * LANES);
        f32x2_t sa = pack2(0.f, 0.f), sb = sa;
        const ulonglong2 z = make_ulonglong2(0ull, 0ull);
#pragma unroll
        for (int st = 0; st < LANES / 2; ++st) {
          const int cp = (st + lane) & (LANES / 2 - 1);
          const ulonglong2 v = row[cp];
          row[cp] = z;
          sa = add2(sa, v.x);
          sb = add2(sb, v.y);
        }
        float lo, hi;
        unpack2(add2(sa, sb), lo, hi);
        accp[j][0] += (double)lo;
        accp[j][1] += (double)hi;
      }
    }
#pragma unroll
    for (int j = 0; j < NJS; ++j) {
      const int sg = lane + 32 * j;
      if (sg < k * nsingle) {
        float4* row = reinterpret_cast<float4*>(lanesingle + sg * LANES);
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int st = 0; st < LANES / 4; ++st) {
          const int cp = (st + lane) & (LANES / 4 - 1);
          const float4 v = row[cp];
          row[cp] = make_float4(0.f, 0.f, 0.f, 0.f);
          sa += v.x + v.y;
          sb += v.z + v.w;
        }
        accs[j] += (double)(sa + sb);
      }
    }
    __syncwarp();
  };

  bool final_pass = false;      // whole-fit mode: the label-writing scan after the last update
  for (int it = 0;; ++it) {
    // ---- centroid norms in torch's summation order (cs holds this pass's centroids) ----
    __syncthreads();
    if (tid == 0) nan_centroid = 0;
    __syncthreads();
    if (tid < kpad) {
      float bnorm = INFINITY;           // padding columns: similarity -inf
      if (tid < k) {
        float v[DMAX];
#pragma unroll
        for (int i = 0; i < DMAX; ++i) v[i] = (i < d) ? cs[i * KMAX + tid] : 0.f;
        bnorm = sumsq_torch_order<DMAX>(v, d, col_is_sequential(tid, k));
        // NaN / inf / huge centroid: take the torch.max-exact scalar path.  Below KM_FAST_NORM_MAX for both norms
        // |2 dot| <= 2 sqrt(|a|^2 |b|^2) cannot overflow, so ptxas contracting fl(2 dot) - |a|^2 into one FFMA2 in the
        // packed scan (2 dot is exact) yields the same bits as the reference's separate mul_ and sub_.
        if (!(bnorm <= KM_FAST_NORM_MAX)) nan_centroid = 1;
      }
      bn[tid] = bnorm;
      nbn[tid] = -bnorm;
    }
    __syncthreads();
    const bool nan_possible = nan_centroid != 0;
    const bool acc_pass = accumulate && !final_pass;
    int64_t* labels_out = whole_fit ? (final_pass ? fit.labels_final : nullptr) : labels;
    float* maxsims_out = whole_fit ? nullptr : maxsims;
#pragma unroll
    for (int j = 0; j < NJP; ++j) accp[j][0] = accp[j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NJS; ++j) accs[j] = 0.0;
    sim_acc = 0.0;

    constexpr int PPL = 2;                                   // points per lane and iteration (instruction-level parallelism)
    // 32-bit point indices (km_check keeps N + one grid stride below 2^32): one 32-bit compare per point instead of
    // compare pairs; the compile-time-d kernels are only dispatched while d * N < 2^32, so the element offset r * N + i
    // is 32-bit arithmetic too and a load costs one IMAD + one IMAD.WIDE (no 64-bit address chains)
    const uint32_t n32 = (uint32_t)n;
    const uint32_t stride = gridDim.x * (uint32_t)(WARPS * 32 * PPL);
    const int64_t n_all = fit.n_global > 0 ? fit.n_global : n;
    const uint32_t goff = (uint32_t)fit.col_offset;                           // km_check keeps global indices below 2^32
    const uint32_t seq_limit = (n_all >= 4 && n_all < 8) ? 4u : 32u * (uint32_t)(n_all / 32);   // col_is_sequential(g, n_all) == g < seq_limit
    const float* dl_pin = dl;
    asm volatile("" : "+l"(dl_pin));     // opaque: keeps the batch entry's base in a register pair instead of re-deriving
                                         // data + l * d * n (64-bit multiplies) in front of every load
    auto load_coord = [&](int r, uint32_t i) -> float {
      if constexpr (EXACT) return __ldg(dl_pin + (uint32_t)((uint32_t)r * n32 + i));
      else return __ldg(dl + (int64_t)r * n + i);
    };
    int64_t* lab_out = labels_out ? labels_out + (int64_t)l * n : nullptr;
    float* sim_out = maxsims_out ? maxsims_out + (int64_t)l * n : nullptr;
    const int64_t* lab_in = labels_in ? labels_in + (int64_t)l * n : nullptr;
    int since_flush = 0;
    const bool any_out = lab_out != nullptr || sim_out != nullptr;
    // all lanes of a warp iterate together (the loop bound is warp-uniform)
    const uint32_t base0 = ((uint32_t)blockIdx.x * WARPS + warp) * (32 * PPL);
    float a_next[PPL][DMAX];
#pragma unroll
    for (int u = 0; u < PPL; ++u) {
      const uint32_t i0 = base0 + u * 32 + lane;
#pragma unroll
      for (int r = 0; r < DMAX; ++r) a_next[u][r] = ((EXACT || r < d) && i0 < n32) ? load_coord(r, i0) : 0.f;
    }
    for (uint32_t base = base0; base < n32; base += stride) {
      float a[PPL][DMAX];
      uint32_t idx[PPL];
#pragma unroll
      for (int u = 0; u < PPL; ++u) {
        idx[u] = base + u * 32 + lane;
#pragma unroll
        for (int r = 0; r < DMAX; ++r) a[u][r] = a_next[u][r];
        // software prefetch of the next batch: its loads are in flight while this one is scored
        const uint32_t in = idx[u] + stride;
#pragma unroll
        for (int r = 0; r < DMAX; ++r) a_next[u][r] = ((EXACT || r < d) && in < n32) ? load_coord(r, in) : 0.f;
      }
      float best[PPL];
      int label[PPL];
      if (lab_in) {   // compute_centroids with caller-supplied labels: accumulation only
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          best[u] = 0.f;
          const int64_t li = idx[u] < n32 ? lab_in[idx[u]] : -1;
          label[u] = (li >= 0 && li < k) ? (int)li : -1;
        }
      } else {
        float anorm[PPL];
        bool finite = !nan_possible;
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
          anorm[u] = sumsq_torch_order<DMAX>(a[u], d, goff + idx[u] < seq_limit);
          finite = finite && (fabsf(anorm[u]) <= KM_FAST_NORM_MAX);   // finite (and not huge) iff every coordinate is
        }
        if (finite) {
          best_centroid_packed<DMAX, EXACT, PPL, KPAD, true, TOURN>(a, d, anorm, cs, KMAX, nbn, kpad, best, label);
        } else {
#pragma unroll
          for (int u = 0; u < PPL; ++u)
            best_centroid<DMAX, EXACT, true>(a[u], d, anorm[u], cs, KMAX, bn, k, kpad, best[u], label[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < PPL; ++u) {
        const bool live = idx[u] < n32;
        if (any_out && live) {                            // (any_out is warp-uniform: accumulation passes skip this block)
          if (lab_out) lab_out[idx[u]] = label[u];
          if (sim_out) sim_out[idx[u]] = best[u];
        }
        if (acc_pass) {                                   // (warp-uniform)
          const bool upd = live && label[u] >= 0;
#pragma unroll
          for (int h = 0; h < SHARE; ++h) {               // SHARE == 2: the half-warps take turns on the shared columns
            if (upd && (SHARE == 1 || (lane >> 4) == h)) {
              const int col = lane & (LANES - 1);
              f32x2_t* pslot = reinterpret_cast<f32x2_t*>(lanerec) + (label[u] * npair) * LANES + col;
#pragma unroll
              for (int p2 = 0; p2 < DMAX / 2; ++p2)
                if (EXACT || p2 < npair) pslot[p2 * LANES] = add2(pslot[p2 * LANES], pack2(a[u][2 * p2], a[u][2 * p2 + 1]));
              float* sslot = lanesingle + (label[u] * nsingle) * LANES + col;
              if (d & 1) {
                float last = 0.f;
#pragma unroll
                for (int r = 0; r < DMAX; ++r)
                  if (r == d - 1) last = a[u][r];
                sslot[0] += last;
                sslot[LANES] += 1.0f;
              } else {
                sslot[0] += 1.0f;
              }
            }
            if (SHARE > 1) __syncwarp();
          }
          if (upd) sim32 += best[u];
        }
      }
      if (acc_pass && (since_flush += PPL) >= KM_FLUSH_EVERY) {
        flush();
        since_flush = 0;
      }
    }
    if (!acc_pass) break;
    flush();

    // ---- block reduction in fixed warp order: every warp parks its folded record in its own (now all-zero) lane-record
    // area, one barrier, then thread e adds entry e over the warps in ascending order (the same sums, bit for bit, as
    // letting the warps add into blk[] one after the other -- without the WARPS serial rounds) ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sim_acc += __shfl_xor_sync(0xffffffffu, sim_acc, o);
    {
      double* wst = reinterpret_cast<double*>(lanerec);
#pragma unroll
      for (int j = 0; j < NJP; ++j) {
        const int rho = lane + 32 * j;
        if (rho < k * npair) {
          const int c = rho / npair, e = c * (d + 1) + 2 * (rho - c * npair);
          wst[e] = accp[j][0];
          wst[e + 1] = accp[j][1];
        }
      }
#pragma unroll
      for (int j = 0; j < NJS; ++j) {
        const int sg = lane + 32 * j;
        if (sg < k * nsingle) {
          const int c = sg / nsingle;
          wst[c * (d + 1) + 2 * npair + (sg - c * nsingle)] = accs[j];
        }
      }
      if (lane == 0) wst[rec] = sim_acc;
    }
    __syncthreads();
    {
      const double* st0 = reinterpret_cast<const double*>(lanerec - (size_t)warp * rec * LANES);   // warp 0's area
      for (int e = tid; e <= rec; e += THREADS) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) sum += st0[(size_t)w * rec * (LANES / 2) + e];
        blk[e] = sum;
      }
    }
    __syncthreads();
    for (int e = lane; e < 2 * (rec + 1); e += 32) lanerec[e] = 0.f;        // the staging words are lane records again
    // partial records are stored entry-major ([l][entry][block]) so that the fold below reads contiguous doubles
    // (whole-fit mode alternates between two partial arrays: a block that is already writing the partials of iteration
    // it + 1 must not disturb a slower block that still folds those of iteration it -- only ONE grid barrier separates them)
    double* lpart = partials + (whole_fit ? (size_t)(it & 1) * fit.part_stride : 0) + (size_t)l * gridDim.x * out_rec;
    for (int e = tid; e < out_rec; e += THREADS) lpart[(size_t)e * gridDim.x + blockIdx.x] = blk[e];
    if (!whole_fit) {
      // ---- grid fold: every warp of this batch entry's blocks sums a few record entries over its blocks' partials ----
      grid_barrier(barrier_ctr, nblocks);
      for (int e = blockIdx.x * WARPS + warp; e < out_rec; e += gridDim.x * WARPS) {
        const double tot = warp_fold_contig(lpart + (size_t)e * gridDim.x, (int)gridDim.x, lane);
        if (lane == 0) {
          if (e < rec) {
            const int c = e / (d + 1), r = e % (d + 1);
            if (r < d) sums[((int64_t)l * d + r) * k + c] += tot;
            else counts[(int64_t)l * k + c] += tot;
          } else if (simsum) {
            simsum[l] += tot;
          }
        }
      }
      break;
    }

    // ---- whole-fit mode: ONE grid barrier per iteration.  Behind it the blocks fold the record entries in groups of
    // KM_FOLD_GROUP: every member folds its share of the entries over ALL block partials (the same warp_fold_contig over
    // the same entry-major partials as the launch-per-iteration path, hence the same bits), writes them to the group's
    // record, the group synchronises among its few members, and every member reads the complete record.  Every group
    // computes the same totals, so no result crosses groups and no second GRID barrier is needed; compared with every
    // block folding everything (148 x 141 x 148 doubles = 24.7 MB of L2 reads per iteration, 4.4 us) the fold traffic
    // drops by the group size.  Then division, error and convergence test, identically in every block
    // (compute_centroids' division kmeans.py:183, calculate_error kmeans.py:45-51, `if error <= self.tol: break` :239).
    km_barrier(barrier_ctr + 2, nblocks, ++phase);
    {
      const int gid = blockIdx.x / KM_FOLD_GROUP, g0 = gid * KM_FOLD_GROUP;
      const int gsz = (int)gridDim.x - g0 < KM_FOLD_GROUP ? (int)gridDim.x - g0 : KM_FOLD_GROUP;
      const int ngroups = ((int)gridDim.x + KM_FOLD_GROUP - 1) / KM_FOLD_GROUP;
      const int member = blockIdx.x - g0;
      const int per_member = (out_rec + gsz - 1) / gsz;
      const int m_lo = member * per_member, m_hi = m_lo + per_member < out_rec ? m_lo + per_member : out_rec;
      double* grec = fit.gtot + ((size_t)l * ngroups + gid) * out_rec;
      constexpr int NE = 4;                       // entries folded together (their loads overlap)
      const int per_warp = (m_hi - m_lo + WARPS - 1) / WARPS;
      const int e_lo = m_lo + warp * per_warp, e_hi = e_lo + per_warp < m_hi ? e_lo + per_warp : m_hi;
      for (int e = e_lo; e < e_hi; e += NE) {
        double tot[NE];
        const int valid = e_hi - e < NE ? e_hi - e : NE;
        warp_fold_contig_multi<NE>(lpart + (size_t)e * gridDim.x, gridDim.x, valid, (int)gridDim.x, lane, tot);
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < NE; ++u)
            if (u < valid) grec[e + u] = tot[u];
        }
      }
      if (gsz > 1) km_barrier(fit.gctr + (size_t)l * ngroups + gid, (unsigned)gsz, gphase);      // the group's members only
      else __syncthreads();
      ++gphase;
      for (int e = tid; e < out_rec; e += THREADS) blk[e] = __ldcg(grec + e);
    }
    __syncthreads();
    const bool sharded = fit.world > 1;
    const size_t slot = (size_t)gridDim.y * out_rec;
    if (sharded) {
      // Row-sharded fit: block 0 of every batch entry stores this rank's folded record straight into every rank's
      // exchange slot (its own included) as stamped packets; every block then gathers the `world` records from its OWN
      // buffer, entry by entry, spinning on a packet until it carries this iteration's stamp, and adds them in rank
      // order -- so all ranks hold bit-identical totals, centroids and convergence decisions without a broadcast.
      const unsigned stamp = fit.stamp_base + 1u + (unsigned)it;
      const size_t slot_words = (size_t)gridDim.y * out_rec;
      if (blockIdx.x == 0) {
        if (it == 0) {           // peers must have entered this call (they are done reading the previous call's slots)
          if (tid == 0) xchg_wait_all(xchg_ready(fit.xchg[fit.rank]), fit.world, fit.stamp_base);
          __syncthreads();
        }
        for (int e = tid; e < out_rec; e += THREADS) {
          const unsigned long long v = (unsigned long long)__double_as_longlong(blk[e]);
          for (int p = 0; p < fit.world; ++p)
            ll_store(xchg_packets(fit.xchg[p], it & 1, fit.world, fit.rank, slot_words, (size_t)l * out_rec + e), v, stamp);
        }
      }
      __syncthreads();           // blk[] is overwritten with the global totals below
      for (int e = tid; e < out_rec; e += THREADS) {
        double t = 0.0;
        for (int r = 0; r < fit.world; ++r)
          t += __longlong_as_double((long long)ll_load(
              xchg_packets(fit.xchg[fit.rank], it & 1, fit.world, r, slot_words, (size_t)l * out_rec + e), stamp));
        blk[e] = t;
      }
      __syncthreads();
    }
    // folded total of record entry idx of this batch entry (this rank's fold, or the ranks' records added in rank order)
    auto total_at = [&](int idx) -> double { return blk[idx]; };
    double e2 = 0.0;
    for (int e = tid; e < d * k; e += THREADS) {
      const int r = e / k, c = e - r * k;
      const float v = (float)(total_at(c * (d + 1) + r) / total_at(c * (d + 1) + d));   // 0/0 -> NaN as the reference
      csn[r * KMAX + c] = v;
      const double df = (double)cs[r * KMAX + c] - (double)v;
      e2 += df * df;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    if (lane == 0) red_s[warp] = e2;
    __syncthreads();
    double err = 0.0;
    for (int w = 0; w < WARPS; ++w) err += red_s[w];
    if (gridDim.y > 1) {       // the reference's error runs over every batch entry: exchange the per-entry errors
      double* errs = fit.totals + (size_t)gridDim.y * out_rec;
      if (blockIdx.x == 0 && tid == 0) errs[l] = err;
      km_barrier(barrier_ctr + 2, nblocks, ++phase);
      err = 0.0;
      for (int y = 0; y < (int)gridDim.y; ++y) err += __ldcg(errs + y);
    }
    const bool converged = err <= fit.tol;
    final_pass = converged || it + 1 >= fit.max_iter;
    const bool stop = final_pass && !fit.labels_final;     // nobody wants the labels: skip the last scan
    if (blockIdx.x == 0) {
      if (final_pass) {
        for (int e = tid; e < d * k; e += THREADS) fit.cent_out[(int64_t)l * d * k + e] = csn[(e / k) * KMAX + (e % k)];
        if (tid == 0 && fit.simsum_last) fit.simsum_last[l] = total_at(rec);
      }
      if (l == 0 && tid == 0 && final_pass) {
        if (fit.err) fit.err[0] = err;
        if (fit.status) { fit.status[0] = converged ? 1 : 0; fit.status[1] = it + 1; }
      }
    }
    if (!final_pass) {         // next assignment runs against the updated centroids
      __syncthreads();
      for (int e = tid; e < DMAX * KMAX; e += THREADS) cs[e] = csn[e];
    }
    // on the final pass cs keeps the centroids of the last assignment: its labels are the ones fit() returns
    if (stop) break;
  }
  if (whole_fit) km_barrier_exit(barrier_ctr + 2, nblocks);
}

template <int DMAX, int KMAX>
static size_t km_smem_bytes(int d, int k, int warps, bool accumulate, int share = 1) {
  return (size_t)(2 * DMAX * KMAX + 2 * KMAX) * sizeof(float) + (size_t)((KMAX * (DMAX + 1) + 2) & ~1) * sizeof(double) +
         (accumulate ? (size_t)warps * k * (d + 1) * (32 / share) * sizeof(float) : 0);
}

// new = float(sums / counts); err = sum (old - new)^2; clears the accumulators for the next iteration.
__global__ void kmeans_finalize_kernel(double* __restrict__ sums, double* __restrict__ counts, int l, int d, int k,
                                       const float* __restrict__ old_c, float* __restrict__ new_c,
                                       double* __restrict__ err, double tol, int32_t* __restrict__ status,
                                       double* __restrict__ simsum, double* __restrict__ simsum_last) {
  if (status && status[0] != 0) return;
  __shared__ double red[256];
  const int total = l * d * k;
  double e2 = 0.0;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e % k, li = e / (d * k);
    const double cnt = counts[li * k + c];
    const float v = (float)(sums[e] / cnt);
    new_c[e] = v;
    if (old_c) {
      const double df = (double)old_c[e] - (double)v;
      e2 += df * df;
    }
  }
  red[threadIdx.x] = e2;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  // every thread has read what it needs from sums / counts before anyone clears them
  for (int e = threadIdx.x; e < total; e += blockDim.x) sums[e] = 0.0;
  for (int e = threadIdx.x; e < l * k; e += blockDim.x) counts[e] = 0.0;
  if (simsum)
    for (int e = threadIdx.x; e < l; e += blockDim.x) {
      if (simsum_last) simsum_last[e] = simsum[e];   // sum of best similarities of the iteration just finished
      simsum[e] = 0.0;
    }
  if (threadIdx.x == 0) {
    if (err) err[0] = red[0];
    if (status) {
      status[1] += 1;
      if (red[0] <= tol) status[0] = 1;
    }
  }
}

// ---- farthest-point seeding (kmeans.py:78-112) ------------------------------------------------
__device__ __forceinline__ unsigned long long pack_min_key(float v, int64_t idx) {
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone: smaller float -> smaller key
  return ((unsigned long long)u << 32) | (unsigned long long)(uint32_t)idx;
}

// scratch = l*K candidate keys (column 0 = the caller's first index, the others "none yet") followed by 16 words for the
// barrier counters of the persistent kernel, zeroed here on every call (a kernel that died half way cannot poison the next)
constexpr int SEED_SCRATCH_EXTRA = 16;
__global__ void kmeans_seed_init_kernel(unsigned long long* scratch, int l, int k, int64_t first_index) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < l * k) scratch[e] = (e % k == 0) ? (unsigned long long)first_index : ~0ull;
  else if (e < l * k + SEED_SCRATCH_EXTRA) scratch[e] = 0ull;
}

// Step `ncols` (1 <= ncols < K): the chosen points scratch[l*K + 0 .. ncols) are the current centroids; every
// point takes its best similarity against them exactly as the reference recomputes it, and the point with the
// lowest best similarity (lowest index on ties) is recorded in scratch[l*K + ncols].
template <int DMAX, int KMAX, bool EXACT>
__global__ void __launch_bounds__(KM_THREADS) kmeans_seed_step_kernel(const float* __restrict__ data, int d, int64_t n,
                                                                      int k, int ncols,
                                                                      unsigned long long* __restrict__ scratch,
                                                                      const float* __restrict__ cent_in,
                                                                      unsigned long long* __restrict__ key_out,
                                                                      int64_t col_offset, int64_t n_all) {
  __shared__ __align__(16) float cs[DMAX * KMAX];
  __shared__ __align__(16) float bn[KMAX];
  __shared__ unsigned long long wmin[KM_WARPS];
  __shared__ int huge_centroid;
  for (int e = threadIdx.x; e < DMAX * KMAX; e += KM_THREADS) cs[e] = 0.f;
  if (threadIdx.x < KMAX) bn[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) huge_centroid = 0;
  __syncthreads();
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* dl = data + (int64_t)l * d * n;
  for (int e = tid; e < d * ncols; e += KM_THREADS) {
    const int r = e / ncols, j = e % ncols;
    if (cent_in) {   // explicit centroids (l,d,K): the row-sharded path, where chosen points may live on another rank
      cs[r * KMAX + j] = __ldg(cent_in + ((int64_t)l * d + r) * k + j);
    } else {
      const int64_t idx = (int64_t)(scratch[(int64_t)l * k + j] & 0xffffffffull);
      cs[r * KMAX + j] = __ldg(dl + (int64_t)r * n + idx);
    }
  }
  __syncthreads();
  const int kpad = (ncols + 3) & ~3;
  if (tid < kpad) {
    float bnorm = INFINITY;
    if (tid < ncols) {
      float v[DMAX];
#pragma unroll
      for (int i = 0; i < DMAX; ++i) v[i] = (i < d) ? cs[i * KMAX + tid] : 0.f;
      bnorm = sumsq_torch_order<DMAX>(v, d, col_is_sequential(tid, ncols));
      if (!(bnorm <= KM_FAST_NORM_MAX)) huge_centroid = 1;
    }
    bn[tid] = -bnorm;     // negated for the packed scan; padding columns: similarity -inf
  }
  __syncthreads();
  const bool scalar_only = huge_centroid != 0;
  if (scalar_only) {      // the scalar scan subtracts |b|^2
    __syncthreads();
    if (tid < kpad) bn[tid] = -bn[tid];
    __syncthreads();
  }
  unsigned long long key = ~0ull;
  const int64_t sstride = (int64_t)gridDim.x * KM_THREADS;
  float a_next[DMAX];
  {
    const int64_t i0 = (int64_t)blockIdx.x * KM_THREADS + tid;
#pragma unroll
    for (int r = 0; r < DMAX; ++r) a_next[r] = ((EXACT || r < d) && i0 < n) ? __ldg(dl + (int64_t)r * n + i0) : 0.f;
  }
  for (int64_t i = (int64_t)blockIdx.x * KM_THREADS + tid; i < n; i += sstride) {
    float a[1][DMAX];
#pragma unroll
    for (int r = 0; r < DMAX; ++r) a[0][r] = a_next[r];
    {
      const int64_t in = i + sstride;
#pragma unroll
      for (int r = 0; r < DMAX; ++r) a_next[r] = ((EXACT || r < d) && in < n) ? __ldg(dl + (int64_t)r * n + in) : 0.f;
    }
    const float an[1] = {sumsq_torch_order<DMAX>(a[0], d, col_is_sequential(col_offset + i, n_all))};
    float best[1];
    int label[1];
    if (!scalar_only && fabsf(an[0]) <= KM_FAST_NORM_MAX)
      best_centroid_packed<DMAX, EXACT, 1, 0, false>(a, d, an, cs, KMAX, bn, kpad, best, label);
    else if (scalar_only)
      best_centroid<DMAX, EXACT, false>(a[0], d, an[0], cs, KMAX, bn, ncols, kpad, best[0], label[0]);
    else
      best_centroid_neg<DMAX, EXACT>(a[0], d, an[0], cs, KMAX, bn, kpad, best[0]);
    const unsigned long long kk = pack_min_key(best[0], i);
    key = kk < key ? kk : key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other < key ? other : key;
  }
  if (lane == 0) wmin[warp] = key;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < KM_WARPS; ++w) key = wmin[w] < key ? wmin[w] : key;
    atomicMin(key_out ? &key_out[l] : &scratch[(int64_t)l * k + ncols], key);
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, int count, unsigned long long v) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < count) p[e] = v;
}

// Row-shard seeding helpers.  A local candidate key (order-preserving similarity bits << 32 | local index) becomes a
// GLOBAL key that a signed 64-bit MIN all-reduce orders correctly: the index is shifted by the shard's row offset and
// the top bit is flipped (unsigned order -> signed order); "no candidate" (empty shard) is INT64_MAX.
__global__ void kmeans_seed_globalize_kernel(unsigned long long* __restrict__ key, int l, unsigned long long row_offset) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= l) return;
  const unsigned long long k = key[e];
  key[e] = (k == ~0ull) ? 0x7fffffffffffffffull : ((k + row_offset) ^ 0x8000000000000000ull);
}

// coords (l,d) float64 = coordinates of the winning global column if this shard owns it, else 0 (the caller sums over ranks)
__global__ void kmeans_seed_fetch_kernel(const float* __restrict__ data, int l, int d, int64_t n, int64_t row_offset,
                                         const long long* __restrict__ gkey, double* __restrict__ coords) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= l * d) return;
  const int li = e / d, r = e % d;
  const int64_t gidx = (int64_t)(((unsigned long long)gkey[li] ^ 0x8000000000000000ull) & 0xffffffffull);
  const int64_t local = gidx - row_offset;
  coords[e] = (local >= 0 && local < n) ? (double)__ldg(data + ((int64_t)li * d + r) * n + local) : 0.0;
}

__global__ void kmeans_seed_gather_kernel(const float* __restrict__ data, int l, int d, int64_t n, int k,
                                          const unsigned long long* __restrict__ scratch, float* __restrict__ centroids) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= l * d * k) return;
  const int j = e % k, r = (e / k) % d, li = e / (d * k);
  const int64_t idx = (int64_t)(scratch[(int64_t)li * k + j] & 0xffffffffull);
  centroids[e] = __ldg(data + ((int64_t)li * d + r) * n + idx);
}

// ---- farthest-point seeding, all K - 1 steps in ONE persistent cooperative launch ------------------------------------
// Every block keeps its chunk of points in shared memory (coordinate-major, loaded once) and every thread keeps the
// running best similarity and |a|^2 of its <= SEEDP_MMAX points in registers, so a step costs one scan of the resident
// points against the NEWEST centroid only.  That is exact, not an approximation: the similarity of a point to an
// already chosen centroid changes from one step to the next only when the summation order of the centroid norms
// changes (col_is_sequential depends on the number of columns: at 4, 8 and 32 columns), and at exactly those steps the
// running best is recomputed against all columns -- the reference recomputes everything every step (kmeans.py:95-98)
// and obtains the same bits.  Every similarity is the separately rounded scalar sequence fl(fl(fl(2 dot) - |a|^2) -
// |b|^2), which is what all three paths of kmeans_seed_step_kernel produce.  Per step: block arg-min -> 64-bit
// atomicMin -> ONE grid barrier -> every block reads the winner.  Row shards (world > 1): after the barrier block (0,0)
// stores this rank's candidate (key with GLOBAL index, coordinates) into every rank's exchange slot and raises its flag;
// all ranks then take the smallest key in rank order -- the (value, index) arg-min exchange and the broadcast of the
// winner of SURVEY section 8e, inside the kernel.
constexpr int SEEDP_THREADS = 1024;
constexpr int SEEDP_MMAX = 10;              // resident points per thread
constexpr int SEEDP_SMEM_MAX = 208 * 1024;  // bytes of resident coordinates per block

struct SeedShard {
  int rank, world;
  unsigned char* const* xchg;   // as KmLloyd::xchg
  unsigned stamp_base;          // flags of this call: stamp_base (ready), stamp_base + 1 + step
  int64_t col_offset, n_global; // global numbering of this shard's columns (n_global = 0: numbered on its own)
};

template <int DMAX, int KMAX, bool EXACT>
__global__ void __launch_bounds__(SEEDP_THREADS, 1) kmeans_seed_persistent_kernel(
    const float* __restrict__ data, int d, int64_t n, int k, int64_t first_index, unsigned long long* __restrict__ scratch,
    float* __restrict__ centroids, unsigned* __restrict__ barrier_ctr, int pts_per_block, const SeedShard sh) {
  extern __shared__ __align__(16) float seed_xs[];               // [d][pts_per_block]
  __shared__ float cs[DMAX * KMAX];                               // chosen centroids, [coordinate][column]
  __shared__ float bn[KMAX];                                      // their squared norms under the current column count
  __shared__ unsigned long long wmin[SEEDP_THREADS / 32];
  __shared__ unsigned long long win_s;
  __shared__ float win_xyz[DMAX];
  if (EXACT) d = DMAX;
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = pts_per_block;
  const int64_t p0 = (int64_t)blockIdx.x * P;
  const int np = (int)(n - p0 < P ? (n - p0 > 0 ? n - p0 : 0) : P);
  const float* dl = data + (int64_t)l * d * n;
  const unsigned nblocks = gridDim.x * gridDim.y;
  const bool sharded = sh.world > 1;
  const int64_t n_all = sh.n_global > 0 ? sh.n_global : n;
  const int64_t goff = sh.col_offset + p0;                        // global index of this block's first point
  unsigned phase = 0;
  if (sharded && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
  {
    __threadfence_system();
    for (int p = 0; p < sh.world; ++p) st_relaxed_sys(xchg_ready(sh.xchg[p]) + sh.rank, sh.stamp_base);
  }

  for (int r = 0; r < d; ++r)
    for (int t = tid; t < np; t += SEEDP_THREADS) seed_xs[r * P + t] = __ldg(dl + (int64_t)r * n + p0 + t);
  __syncthreads();
  float best[SEEDP_MMAX], an[SEEDP_MMAX];
#pragma unroll
  for (int m = 0; m < SEEDP_MMAX; ++m) {
    const int t = m * SEEDP_THREADS + tid;
    best[m] = -INFINITY;
    an[m] = 0.f;
    if (t < np) {
      float a[DMAX];
#pragma unroll
      for (int r = 0; r < DMAX; ++r) a[r] = (EXACT || r < d) ? seed_xs[r * P + t] : 0.f;
      an[m] = sumsq_torch_order<DMAX>(a, d, col_is_sequential(goff + t, n_all));
    }
  }

  for (int i = 0; i < k; ++i) {                                    // elect column i
    if (i > 0) {
      // columns whose similarities must be (re)computed this step: all of them when a norm changes its summation order
      // (col_is_sequential(j, ncols) changes for an existing column j exactly when ncols reaches 4, 8 or 32)
      const bool full = i == 1 || i == 4 || i == 8 || i == 32;
      static_assert(ET_MAX_CLUSTERS <= 64, "the regime boundaries above cover up to 64 columns");
      const int jlo = full ? 0 : i - 1;
      if (tid >= jlo && tid < i) {
        float v[DMAX];
#pragma unroll
        for (int r = 0; r < DMAX; ++r) v[r] = (r < d) ? cs[r * KMAX + tid] : 0.f;
        bn[tid] = sumsq_torch_order<DMAX>(v, d, col_is_sequential(tid, i));
      }
      __syncthreads();
      unsigned long long key = ~0ull;
#pragma unroll
      for (int m = 0; m < SEEDP_MMAX; ++m) {
        const int t = m * SEEDP_THREADS + tid;
        if (t < np) {
          float a[DMAX];
#pragma unroll
          for (int r = 0; r < DMAX; ++r) a[r] = (EXACT || r < d) ? seed_xs[r * P + t] : 0.f;
          float b = full ? -INFINITY : best[m];
          for (int j = jlo; j < i; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int r = 0; r < DMAX; ++r)
              if (EXACT || r < d) dot = fmaf(a[r], cs[r * KMAX + j], dot);
            const float y = __fsub_rn(__fsub_rn(__fmul_rn(dot, 2.0f), an[m]), bn[j]);
            if (y > b) b = y;
          }
          best[m] = b;
          const unsigned long long kk = pack_min_key(b, goff + t);
          key = kk < key ? kk : key;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
      }
      if (lane == 0) wmin[warp] = key;
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < SEEDP_THREADS / 32; ++w) key = wmin[w] < key ? wmin[w] : key;
        if (key != ~0ull) atomicMin(&scratch[(int64_t)l * k + i], key);
      }
      km_barrier(barrier_ctr, nblocks, ++phase);
    }
    // ---- the winner of this step (column 0: the caller's first index, stored by the init kernel) ----
    if (sharded) {
      const unsigned stamp = sh.stamp_base + 1u + (unsigned)i;
      constexpr int recd = 1 + (DMAX + 1) / 2;                     // 64-bit words per candidate record: key, coordinate pairs
      const size_t slot_words = (size_t)gridDim.y * ((size_t)k * (d + 1) + 1);     // payload words per rank and parity
      if (blockIdx.x == 0 && blockIdx.y == 0) {
        if (i == 0) {
          if (tid == 0) xchg_wait_all(xchg_ready(sh.xchg[sh.rank]), sh.world, sh.stamp_base);
          __syncthreads();
        }
        if (tid < (int)gridDim.y) {                                // one thread per batch entry
          const int ll = tid;
          unsigned long long lw = __ldcg(&scratch[(int64_t)ll * k + i]);
          const int64_t local = (int64_t)(lw & 0xffffffffull) - sh.col_offset;
          const bool mine = lw != ~0ull && local >= 0 && local < n;
          if (!mine) lw = ~0ull;                                    // (column 0: only the owner proposes it)
          float xyz[DMAX];
#pragma unroll
          for (int r = 0; r < DMAX; ++r) xyz[r] = (mine && r < d) ? __ldg(data + ((int64_t)ll * d + r) * n + local) : 0.f;
          for (int p = 0; p < sh.world; ++p) {
            ll_store(xchg_packets(sh.xchg[p], i & 1, sh.world, sh.rank, slot_words, (size_t)ll * recd), lw, stamp);
#pragma unroll
            for (int r = 0; r < (DMAX + 1) / 2; ++r)
              ll_store(xchg_packets(sh.xchg[p], i & 1, sh.world, sh.rank, slot_words, (size_t)ll * recd + 1 + r),
                       pack2(xyz[2 * r], 2 * r + 1 < DMAX ? xyz[2 * r + 1] : 0.f), stamp);
          }
        }
      }
      if (tid == 0) {              // every block: the smallest key over the ranks (rank order), then its coordinates
        unsigned long long bestk = ~0ull;
        int who = 0;
        for (int r = 0; r < sh.world; ++r) {
          const unsigned long long kr = ll_load(xchg_packets(sh.xchg[sh.rank], i & 1, sh.world, r, slot_words, (size_t)l * recd), stamp);
          if (kr < bestk) { bestk = kr; who = r; }
        }
#pragma unroll
        for (int r = 0; r < (DMAX + 1) / 2; ++r) {
          float lo, hi;
          unpack2(ll_load(xchg_packets(sh.xchg[sh.rank], i & 1, sh.world, who, slot_words, (size_t)l * recd + 1 + r), stamp), lo, hi);
          win_xyz[2 * r] = lo;
          if (2 * r + 1 < DMAX) win_xyz[2 * r + 1] = hi;
        }
        win_s = bestk;
      }
      __syncthreads();
    } else {
      if (tid == 0) win_s = __ldcg(&scratch[(int64_t)l * k + i]);
      __syncthreads();
      const int64_t idx = (int64_t)(win_s & 0xffffffffull);
      if (tid < d) win_xyz[tid] = __ldg(dl + (int64_t)tid * n + idx);
      __syncthreads();
    }
    if (tid < d) {
      const float v = win_xyz[tid];
      cs[tid * KMAX + i] = v;
      if (blockIdx.x == 0) centroids[((int64_t)l * d + tid) * k + i] = v;
    }
    __syncthreads();
  }
  km_barrier_exit(barrier_ctr, nblocks);
}

// Launch the assign kernel (cooperatively when it accumulates: the fold needs a grid barrier).
template <int DMAX, int KMAX, int WARPS, bool EXACT, int KPAD, bool TOURN = false, int SHARE = 1>
static int km_launch_k(const float* data, const float* centroids, int l, int d, int64_t n, int k, int64_t* labels,
                     float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                     const int32_t* status, const int64_t* labels_in, KmLloyd fit, cudaStream_t st) {
  if constexpr (!TOURN && EXACT && KPAD == 20 && WARPS == 12) {     // A/B switch, reference shape only
    if (tune_get(ET_TUNE_KM_TOURNAMENT))
      return km_launch_k<DMAX, KMAX, WARPS, EXACT, KPAD, true>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum,
                                                               workspace, status, labels_in, fit, st);
  }
  auto kern = kmeans_assign_kernel<DMAX, KMAX, WARPS, EXACT, KPAD, TOURN, SHARE>;
  constexpr int KM_THREADS_L = WARPS * 32;
  const bool coop = sums != nullptr || fit.cent_out != nullptr;    // accumulating launches fold over the grid
  const size_t smem = km_smem_bytes<DMAX, KMAX>(d, k, WARPS, coop, SHARE);
  if (smem > 226 * 1024) return fail(ET_ERR_UNSUPPORTED, "k-means: K (d+1) = %d too large for the accumulation records", k * (d + 1));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_assign_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KM_THREADS_L, smem);
  if (e != cudaSuccess || per_sm < 1) return fail(ET_ERR_CUDA, "kmeans_assign_kernel: occupancy query failed");
  if (per_sm > 8) per_sm = 8;
  int64_t cap = (int64_t)sm_count() * per_sm / l;     // all l * grid.x blocks must be co-resident
  if (cap < 1) {
    // more batch entries than co-resident blocks: the whole-fit kernel cannot run (its convergence test spans all
    // entries); single passes are cut into launches of as many entries as fit, one block each
    if (fit.cent_out)
      return fail(ET_ERR_UNSUPPORTED, "k-means: batch l = %d exceeds the co-resident block budget of the whole-fit kernel", l);
    const int lc = sm_count() * per_sm;
    for (int l0 = 0; l0 < l; l0 += lc) {
      const int ll = l - l0 < lc ? l - l0 : lc;
      auto off = [&](auto* p, int64_t stride) { return p ? p + (int64_t)l0 * stride : p; };
      int rc = km_launch_k<DMAX, KMAX, WARPS, EXACT, KPAD, TOURN, SHARE>(off(data, (int64_t)d * n), off(centroids, (int64_t)d * k), ll, d, n, k,
                                                          off(labels, n), off(maxsims, n), off(sums, (int64_t)d * k),
                                                          off(counts, k), off(simsum, 1), workspace, status, off(labels_in, n),
                                                          fit, st);
      if (rc) return rc;
    }
    return ET_OK;
  }
  int64_t gx = (n + 2 * KM_THREADS_L - 1) / (2 * KM_THREADS_L);     // two points per lane and iteration
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)l);
  unsigned* ctr = reinterpret_cast<unsigned*>(workspace);
  double* parts = workspace ? reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 128) : nullptr;
  if (fit.cent_out) {      // whole-fit scratch behind the partial records: totals / errors, group records, group counters
    const size_t rec1 = (size_t)k * (d + 1) + 1, slots = (size_t)sm_count() * 8;
    fit.part_stride = slots * rec1;
    fit.totals = parts + 2 * slots * rec1;
    fit.gtot = fit.totals + (size_t)l * (rec1 + 1);
    fit.gctr = reinterpret_cast<unsigned*>(fit.gtot + slots * rec1);
  }
  if (fit.cent_out) {
    // whole-fit launches start from a clean barrier header whatever an earlier, aborted launch may have left behind
    e = cudaMemsetAsync(workspace, 0, 128, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(fit.gctr, 0, (size_t)sm_count() * 8 * sizeof(unsigned), st);
    if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_assign_kernel: cudaMemsetAsync: %s", cudaGetErrorString(e));
  }
  if (coop) {
    e = launch_cooperative(kern, grid, dim3(KM_THREADS_L), smem, st, data, centroids, d, n, k, labels, maxsims, sums, counts,
                           simsum, ctr, parts, status, labels_in, fit);
    if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_assign_kernel: cooperative launch: %s", cudaGetErrorString(e));
  } else {
    kern<<<grid, KM_THREADS_L, smem, st>>>(data, centroids, d, n, k, labels, maxsims, sums, counts, simsum, ctr, parts, status,
                                         labels_in, fit);
  }
  return check_launch("kmeans_assign_kernel");
}

// 20 anchors (padded count 20) is the reference configuration: its scan is unrolled at compile time
template <int DMAX, int KMAX, int WARPS, bool EXACT>
static int km_launch_w(const float* data, const float* centroids, int l, int d, int64_t n, int k, int64_t* labels,
                       float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                       const int32_t* status, const int64_t* labels_in, const KmLloyd& fit, cudaStream_t st) {
  if (EXACT && ((k + 3) & ~3) == 20)
    return km_launch_k<DMAX, KMAX, WARPS, EXACT, EXACT ? 20 : 0>(data, centroids, l, d, n, k, labels, maxsims, sums, counts,
                                                                 simsum, workspace, status, labels_in, fit, st);
  return km_launch_k<DMAX, KMAX, WARPS, EXACT, 0>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum,
                                                  workspace, status, labels_in, fit, st);
}

template <int DMAX, int KMAX, bool EXACT>
static int km_launch(const float* data, const float* centroids, int l, int d, int64_t n, int k, int64_t* labels,
                     float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                     const int32_t* status, const int64_t* labels_in, const KmLloyd& fit, cudaStream_t st) {
  // accumulating launches of the reference shape: ONE 12-warp block per SM while its lane-private records fit shared
  // memory -- a third of the partial records and grid-barrier arrivals of the 3 x 4-warp configuration
  if constexpr (EXACT) {
    // A/B: more warps per SM on half-size records (two lanes per record column, see SHARE)
    const int wide = tune_get(ET_TUNE_KM_WARPS);
    if (wide && (sums != nullptr || fit.cent_out != nullptr) && l <= sm_count() && ((k + 3) & ~3) == 20) {
      if (wide == 16)
        return km_launch_k<DMAX, KMAX, 16, EXACT, 20, false, 2>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum,
                                                                workspace, status, labels_in, fit, st);
    }
    if ((sums != nullptr || fit.cent_out != nullptr) && l <= sm_count() && km_smem_bytes<DMAX, KMAX>(d, k, 12, true) <= 224 * 1024)
      return km_launch_w<DMAX, KMAX, 12, EXACT>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum, workspace,
                                                status, labels_in, fit, st);
  }
  // four warps per block while their lane-private records fit ~100 KB, otherwise one warp per block
  if (km_smem_bytes<DMAX, KMAX>(d, k, KM_WARPS, sums != nullptr || fit.cent_out != nullptr) <= 100 * 1024)
    return km_launch_w<DMAX, KMAX, KM_WARPS, EXACT>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum,
                                                    workspace, status, labels_in, fit, st);
  return km_launch_w<DMAX, KMAX, 1, EXACT>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum, workspace,
                                           status, labels_in, fit, st);
}

// (6, <=32) is the reference configuration (k = 6 coefficients, 20 anchors): compile-time d.
static int km_dispatch(const float* data, const float* centroids, int l, int d, int64_t n, int k, int64_t* labels,
                       float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                       const int32_t* status, const int64_t* labels_in, const KmLloyd& fit, cudaStream_t st) {
  if (d == 6 && k <= 32 && 6 * n < ((int64_t)1 << 32))      // compile-time d with 32-bit element offsets
    return km_launch<6, 32, true>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum, workspace, status,
                                  labels_in, fit, st);
  if (d <= 8 && k <= 32)
    return km_launch<8, 32, false>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum, workspace, status,
                                   labels_in, fit, st);
  return km_launch<ET_MAX_KM_DIM, ET_MAX_CLUSTERS, false>(data, centroids, l, d, n, k, labels, maxsims, sums, counts, simsum,
                                                          workspace, status, labels_in, fit, st);
}

static int km_grid(int64_t n) {
  int64_t g = (n + KM_THREADS - 1) / KM_THREADS;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

static void launch_seed_step(dim3 grid, cudaStream_t st, const float* data, int d, int64_t n, int k, int ncols,
                             unsigned long long* scratch, const float* cent_in, unsigned long long* key_out,
                             int64_t col_offset = 0, int64_t n_all = 0) {
  if (n_all <= 0) n_all = n;      // the launch holds all columns
  if (d == 6 && k <= 32)
    kmeans_seed_step_kernel<6, 32, true><<<grid, KM_THREADS, 0, st>>>(data, d, n, k, ncols, scratch, cent_in, key_out, col_offset,
                                                                      n_all);
  else if (d <= 8 && k <= 32)
    kmeans_seed_step_kernel<8, 32, false><<<grid, KM_THREADS, 0, st>>>(data, d, n, k, ncols, scratch, cent_in, key_out, col_offset,
                                                                       n_all);
  else
    kmeans_seed_step_kernel<ET_MAX_KM_DIM, ET_MAX_CLUSTERS, false><<<grid, KM_THREADS, 0, st>>>(data, d, n, k, ncols, scratch,
                                                                                               cent_in, key_out, col_offset, n_all);
}

// ---- k-means++ seeding by D^2 sampling with greedy local trials, all K - 1 steps in ONE persistent launch -------------
// What sklearn's KMeans(init="k-means++") does for ETAnchor.anchor_generation (anchor.py:65-71): the next centre is
// drawn with probability proportional to the squared distance to the nearest centre chosen so far; `trials` candidates
// are drawn per step and the one that lowers the total potential most is kept.  The random numbers come from the host
// (uniform[l][K][trials] in [0, 1): [.,0,0] picks the first centre, [.,i,j] the j-th candidate of step i), everything
// else runs here, deterministically: points resident in shared memory (a contiguous run of M points per thread, so the
// cumulative sum that the sampling inverts runs in point order), the running D^2 in registers, all sums in float64 in
// a fixed order.  Per step: block sums of D^2 -> grid barrier -> every block derives the total and its own offset, the
// owners of the `trials` thresholds locate their candidates -> barrier -> potentials of the candidates -> barrier ->
// every block picks the best candidate and updates its D^2.  (tests/ hold a numpy restatement of this algorithm.)
constexpr int SEEDD_THREADS = 1024;
constexpr int SEEDD_MMAX = 9;               // resident points per thread (odd: the strided shared-memory reads are conflict-free)
constexpr int SEEDD_TRIALS_MAX = 8;

template <int DMAX, int KMAX, bool EXACT>
__global__ void __launch_bounds__(SEEDD_THREADS, 1) kmeans_seed_d2_kernel(
    const float* __restrict__ data, int d, int64_t n, int k, int trials, const double* __restrict__ uniform,
    float* __restrict__ centroids, unsigned* __restrict__ barrier_ctr, double* __restrict__ partial /* [l][grid.x][trials + 1] */,
    long long* __restrict__ cand /* [l][trials] */, int pts_per_block) {
  extern __shared__ __align__(16) float seedd_xs[];             // [d][pts_per_block]
  __shared__ double wsum[SEEDD_THREADS / 32][SEEDD_TRIALS_MAX];
  __shared__ double bvals[SEEDD_TRIALS_MAX + 2];
  __shared__ float cxyz[SEEDD_TRIALS_MAX][DMAX];
  if (EXACT) d = DMAX;
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = pts_per_block, M = SEEDD_MMAX;
  const int64_t p0 = (int64_t)blockIdx.x * P;
  const int np = (int)(n - p0 < P ? (n - p0 > 0 ? n - p0 : 0) : P);
  const float* dl = data + (int64_t)l * d * n;
  const double* ul = uniform + (size_t)l * k * trials;
  // block partials: the D^2 sums (phase A) and the candidate potentials (phase C) live in separate arrays -- a block that
  // is already in phase A of the next step must not overwrite what a slower block still reads in phase D of this one
  double* sl = partial + (size_t)l * gridDim.x * (trials + 1);       // [grid.x] sums of D^2
  double* pl = sl + gridDim.x;                                        // [grid.x][trials] potentials
  long long* cl = cand + (size_t)l * trials;
  const unsigned nblocks = gridDim.x * gridDim.y;
  unsigned phase = 0;
  for (int r = 0; r < d; ++r)
    for (int t = tid; t < np; t += SEEDD_THREADS) seedd_xs[r * P + t] = __ldg(dl + (int64_t)r * n + p0 + t);
  __syncthreads();

  auto dist2 = [&](int t, const float* c) -> float {          // squared distance of resident point t to c[0..d)
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < DMAX; ++r)
      if (EXACT || r < d) {
        const float df = seedd_xs[r * P + t] - c[r];
        acc = fmaf(df, df, acc);
      }
    return acc;
  };
  // block-wide sums of up to `cnt` doubles per thread, fixed order (xor tree inside a warp, then warps ascending)
  auto block_sums = [&](double (&v)[SEEDD_TRIALS_MAX], int cnt) {
    for (int j = 0; j < cnt; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
      if (lane == 0) wsum[warp][j] = v[j];
    }
    __syncthreads();
    if (tid < cnt) {
      double sacc = 0.0;
      for (int w = 0; w < SEEDD_THREADS / 32; ++w) sacc += wsum[w][tid];
      bvals[tid] = sacc;
    }
    __syncthreads();
  };
  auto load_centre = [&](int slot, int64_t gidx) {            // coordinates of global column gidx -> cxyz[slot]
    if (tid < d) cxyz[slot][tid] = __ldg(dl + (int64_t)tid * n + gidx);
  };

  // centre 0: the point the first random number selects
  int64_t first = (int64_t)(ul[0] * (double)n);
  if (first > n - 1) first = n - 1;
  if (first < 0) first = 0;
  load_centre(0, first);
  __syncthreads();
  float d2[SEEDD_MMAX];
#pragma unroll
  for (int m = 0; m < SEEDD_MMAX; ++m) {
    const int t = tid * M + m;
    d2[m] = t < np ? dist2(t, cxyz[0]) : 0.f;
  }
  if (blockIdx.x == 0 && tid < d) centroids[((int64_t)l * d + tid) * k] = cxyz[0][tid];
  __syncthreads();

  for (int i = 1; i < k; ++i) {
    // ---- A: block sum of D^2 ----
    double v[SEEDD_TRIALS_MAX];
    double mine = 0.0;
#pragma unroll
    for (int m = 0; m < SEEDD_MMAX; ++m) mine += (double)d2[m];
    v[0] = mine;
    block_sums(v, 1);
    const double bsum = bvals[0];
    if (tid == 0) sl[blockIdx.x] = bsum;
    km_barrier(barrier_ctr, nblocks, ++phase);
    // ---- B: total, this block's offset, and the candidates whose thresholds fall into this block ----
    double offset = 0.0, total = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) {
      const double pb = __ldcg(sl + b);
      if (b < (int)blockIdx.x) offset += pb;
      total += pb;
    }
    // exclusive prefix of the per-thread sums in thread order (= point order)
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    __syncthreads();
    if (lane == 31) wsum[warp][0] = incl;
    __syncthreads();
    double wpre = 0.0;
    for (int w = 0; w < warp; ++w) wpre += wsum[w][0];
    const double tpre = offset + wpre + (incl - mine);          // cumulative D^2 before this thread's first point
    const bool last_block = blockIdx.x == gridDim.x - 1;
    const bool last_thread = (int64_t)(tid + 1) * M >= np && (int64_t)tid * M < np;   // owns the block's last resident point
    for (int j = 0; j < trials; ++j) {
      const double thr = ul[(size_t)i * trials + j] * total;
      // first point whose inclusive cumulative sum exceeds thr; the very last point catches rounding at the top end
      const bool in_range = (thr >= tpre && thr < tpre + mine) || (last_block && last_thread && thr >= tpre + mine);
      if (in_range && (int64_t)tid * M < np) {
        double acc = tpre;
        int pick = -1;
#pragma unroll
        for (int m = 0; m < SEEDD_MMAX; ++m) {
          const int t = tid * M + m;
          acc += (double)d2[m];
          if (pick < 0 && t < np && acc > thr) pick = t;
        }
        if (pick < 0) pick = (np - 1 < tid * M + M - 1) ? np - 1 : tid * M + M - 1;
        cl[j] = p0 + pick;
      }
    }
    if (total <= 0.0 && blockIdx.x == 0 && tid == 0)            // every point coincides with a centre: take point 0
      for (int j = 0; j < trials; ++j) cl[j] = 0;
    km_barrier(barrier_ctr, nblocks, ++phase);
    // ---- C: potential of every candidate ----
    for (int j = 0; j < trials; ++j) load_centre(j, (int64_t)__ldcg(cl + j));
    __syncthreads();
    int best = 0;
    if (trials > 1) {
      for (int j = 0; j < trials; ++j) v[j] = 0.0;
#pragma unroll
      for (int m = 0; m < SEEDD_MMAX; ++m) {
        const int t = tid * M + m;
        if (t < np)
          for (int j = 0; j < trials; ++j) v[j] += (double)fminf(d2[m], dist2(t, cxyz[j]));
      }
      block_sums(v, trials);
      if (tid < trials) pl[blockIdx.x * trials + tid] = bvals[tid];
      km_barrier(barrier_ctr, nblocks, ++phase);
      // ---- D: the candidate with the lowest potential (lowest trial index on ties) ----
      double best_pot = 0.0;
      for (int j = 0; j < trials; ++j) {
        double pot = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) pot += __ldcg(pl + (size_t)b * trials + j);
        if (j == 0 || pot < best_pot) { best_pot = pot; best = j; }
      }
    }
#pragma unroll
    for (int m = 0; m < SEEDD_MMAX; ++m) {
      const int t = tid * M + m;
      if (t < np) d2[m] = fminf(d2[m], dist2(t, cxyz[best]));
    }
    if (blockIdx.x == 0 && tid < d) centroids[((int64_t)l * d + tid) * k + i] = cxyz[best][tid];
    // (sl[] is next written after this step's last barrier and read before the next step's second one; pl[] and the
    // candidate slots are written only behind the next step's first barrier: no block can still be reading them)
    __syncthreads();
  }
  km_barrier_exit(barrier_ctr, nblocks);
}

// Launch the persistent seeding kernel if the points fit the blocks' shared memory and registers; returns 1 when the
// caller has to take the launch-per-step path instead.
template <int DMAX, int KMAX, bool EXACT>
static int seed_persistent_try(const float* data, int l, int d, int64_t n, int k, int64_t first_index,
                               unsigned long long* scratch, float* centroids, const SeedShard& sh, cudaStream_t st) {
  if (l > sm_count() || tune_get(ET_TUNE_SEED_STEPWISE)) return 1;
  const int gx = sm_count() / l;
  int64_t P = (n + gx - 1) / gx;
  P = (P + 31) & ~(int64_t)31;
  if (P < 32) P = 32;
  const size_t smem = (size_t)P * d * sizeof(float);
  if (P > (int64_t)SEEDP_THREADS * SEEDP_MMAX || smem > (size_t)SEEDP_SMEM_MAX) return 1;
  if (sh.world > 1 && (int64_t)k * (d + 1) + 1 < 1 + (DMAX + 1) / 2) return 1;     // candidate record must fit a slot
  auto kern = kmeans_seed_persistent_kernel<DMAX, KMAX, EXACT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_seed_persistent_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int blocks_x = (int)((n + P - 1) / P) > 0 ? (int)((n + P - 1) / P) : 1;
  kmeans_seed_init_kernel<<<(l * k + SEED_SCRATCH_EXTRA + 255) / 256, 256, 0, st>>>(scratch, l, k, first_index);
  int rc = check_launch("kmeans_seed_init_kernel");
  if (rc) return rc;
  unsigned* ctr = reinterpret_cast<unsigned*>(scratch + (size_t)l * k);
  e = launch_cooperative(kern, dim3(blocks_x, l), dim3(SEEDP_THREADS), smem, st, data, d, n, k, first_index, scratch, centroids, ctr,
                         (int)P, sh);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_seed_persistent_kernel: cooperative launch: %s", cudaGetErrorString(e));
  return check_launch("kmeans_seed_persistent_kernel");
}

static int seed_persistent(const float* data, int l, int d, int64_t n, int k, int64_t first_index, unsigned long long* scratch,
                           float* centroids, const SeedShard& sh, cudaStream_t st) {
  if (d == 6 && k <= 32) return seed_persistent_try<6, 32, true>(data, l, d, n, k, first_index, scratch, centroids, sh, st);
  if (d <= 8 && k <= 32) return seed_persistent_try<8, 32, false>(data, l, d, n, k, first_index, scratch, centroids, sh, st);
  return seed_persistent_try<ET_MAX_KM_DIM, ET_MAX_CLUSTERS, false>(data, l, d, n, k, first_index, scratch, centroids, sh, st);
}

template <int DMAX, int KMAX, bool EXACT>
static int seed_d2_launch(const float* data, int l, int d, int64_t n, int k, int trials, const double* uniform, float* centroids,
                          void* workspace, cudaStream_t st) {
  if (l > sm_count()) return fail(ET_ERR_UNSUPPORTED, "et_kmeans_d2_init: batch l = %d exceeds the number of SMs", l);
  const int gx = sm_count() / l;
  int64_t P = (n + gx - 1) / gx;
  P = (P + 31) & ~(int64_t)31;
  if (P < 32) P = 32;
  const size_t smem = (size_t)P * d * sizeof(float);
  if (P > (int64_t)SEEDD_THREADS * SEEDD_MMAX || smem > (size_t)SEEDP_SMEM_MAX)
    return fail(ET_ERR_UNSUPPORTED, "et_kmeans_d2_init: %lld points per batch entry do not fit the resident kernel (at most ~%lld)",
                (long long)n, (long long)gx * ((long long)SEEDP_SMEM_MAX / (d * 4)));
  auto kern = kmeans_seed_d2_kernel<DMAX, KMAX, EXACT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_seed_d2_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  e = cudaMemsetAsync(workspace, 0, 128, st);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_seed_d2_kernel: cudaMemsetAsync: %s", cudaGetErrorString(e));
  const int blocks_x = (int)((n + P - 1) / P) > 0 ? (int)((n + P - 1) / P) : 1;
  unsigned* ctr = reinterpret_cast<unsigned*>(workspace);
  long long* cand = reinterpret_cast<long long*>(reinterpret_cast<char*>(workspace) + 128);
  double* partial = reinterpret_cast<double*>(cand + (size_t)l * trials);
  e = launch_cooperative(kern, dim3(blocks_x, l), dim3(SEEDD_THREADS), smem, st, data, d, n, k, trials, uniform, centroids, ctr,
                         partial, cand, (int)P);
  if (e != cudaSuccess) return fail(ET_ERR_CUDA, "kmeans_seed_d2_kernel: cooperative launch: %s", cudaGetErrorString(e));
  return check_launch("kmeans_seed_d2_kernel");
}

static int km_check(int l, int d, int64_t n, int k) {
  if (l < 1 || l > 65535) return fail(ET_ERR_UNSUPPORTED, "k-means: batch l = %d outside [1, 65535]", l);
  if (d < 1 || d > ET_MAX_KM_DIM) return fail(ET_ERR_UNSUPPORTED, "k-means: d = %d outside [1, %d]", d, ET_MAX_KM_DIM);
  if (k < 1 || k > ET_MAX_CLUSTERS) return fail(ET_ERR_UNSUPPORTED, "k-means: K = %d outside [1, %d]", k, ET_MAX_CLUSTERS);
  if (n < 0 || n >= ((int64_t)1 << 32) - ((int64_t)1 << 21))       // 32-bit point indices + one grid stride must not wrap
    return fail(ET_ERR_UNSUPPORTED, "k-means: N = %lld outside [0, 2^32 - 2^21)", (long long)n);
  return ET_OK;
}

}  // namespace et

using namespace et;

extern "C" {

size_t et_kmeans_workspace_bytes(int l, int d, int k_clusters) {
  if (l < 1 || d < 1 || k_clusters < 1) return 0;
  // 128 B of barrier counters + two alternating arrays of one partial record per co-resident block (at most 8 per SM in
  // total) + the whole-fit scratch of et_kmeans_lloyd (l folded records and l per-entry errors, one record and one
  // counter per fold group)
  const size_t rec = (size_t)k_clusters * (d + 1) + 1, slots = (size_t)sm_count() * 8;
  return 128 + (2 * slots * rec + (size_t)l * (rec + 1) + slots * rec) * sizeof(double) + slots * sizeof(unsigned);
}

int et_kmeans_assign(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters,
                     int64_t* labels, float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                     const int32_t* status, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  if (n == 0) return ET_OK;
  ET_REQUIRE((data && centroids) || n == 0, ET_ERR_BADARG, "et_kmeans_assign: data / centroids null");
  ET_REQUIRE(!sums || (counts && workspace), ET_ERR_BADARG, "et_kmeans_assign: sums given without counts / workspace");
  cudaStream_t st = as_stream(stream);
  return km_dispatch(data, centroids, l, d, n, k_clusters, labels, maxsims, sums, counts, simsum, workspace, status, nullptr,
                     KmLloyd{}, st);
}

int et_kmeans_assign_shard(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters,
                           int64_t* labels, float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                           const int32_t* status, int64_t row_offset, int64_t n_global, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(row_offset >= 0 && n_global >= row_offset + n && km_check(l, d, n_global, k_clusters) == ET_OK, ET_ERR_BADARG,
             "et_kmeans_assign_shard: columns [%lld, %lld) outside a data set of %lld", (long long)row_offset,
             (long long)(row_offset + n), (long long)n_global);
  if (n == 0) return ET_OK;
  ET_REQUIRE(data && centroids, ET_ERR_BADARG, "et_kmeans_assign_shard: data / centroids null");
  ET_REQUIRE(!sums || (counts && workspace), ET_ERR_BADARG, "et_kmeans_assign_shard: sums given without counts / workspace");
  KmLloyd shard{};
  shard.col_offset = row_offset;
  shard.n_global = n_global;
  return km_dispatch(data, centroids, l, d, n, k_clusters, labels, maxsims, sums, counts, simsum, workspace, status, nullptr,
                     shard, as_stream(stream));
}

int et_kmeans_lloyd(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters, int max_iter,
                    double tol, float* centroids_out, int64_t* labels, double* err, int32_t* status, double* simsum_last,
                    void* workspace, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(data && centroids && centroids_out && workspace, ET_ERR_BADARG, "et_kmeans_lloyd: null pointer");
  ET_REQUIRE(n >= 1, ET_ERR_BADARG, "et_kmeans_lloyd: no points");
  ET_REQUIRE(max_iter >= 1, ET_ERR_BADARG, "et_kmeans_lloyd: max_iter = %d < 1", max_iter);
  KmLloyd fit{};
  fit.max_iter = max_iter;
  fit.tol = tol;
  fit.cent_out = centroids_out;
  fit.err = err;
  fit.status = status;
  fit.simsum_last = simsum_last;
  fit.labels_final = labels;
  return km_dispatch(data, centroids, l, d, n, k_clusters, nullptr, nullptr, nullptr, nullptr, nullptr, workspace, nullptr, nullptr,
                     fit, as_stream(stream));
}

size_t et_kmeans_exchange_bytes(int l, int d, int k_clusters, int world) {
  if (l < 1 || d < 1 || k_clusters < 1 || world < 1 || world > KM_XCHG_MAX_WORLD) return 0;
  // header + two parities x world records x l (K (d+1) + 1) payload words x two 8-byte packets per word
  return (size_t)KM_XCHG_HEADER + (size_t)2 * world * l * ((size_t)k_clusters * (d + 1) + 1) * 16;
}

int et_kmeans_lloyd_sharded(const float* data, const float* centroids, int l, int d, int64_t n_local, int k_clusters,
                            int max_iter, double tol, float* centroids_out, int64_t* labels, double* err, int32_t* status,
                            double* simsum_last, void* workspace, int rank, int world, void* const* exchange_peers,
                            unsigned stamp_base, int64_t row_offset, int64_t n_global, et_stream_t stream) {
  int rc = km_check(l, d, n_local, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(centroids && centroids_out && workspace && exchange_peers && (n_local == 0 || data), ET_ERR_BADARG,
             "et_kmeans_lloyd_sharded: null pointer");
  ET_REQUIRE(max_iter >= 1, ET_ERR_BADARG, "et_kmeans_lloyd_sharded: max_iter = %d < 1", max_iter);
  ET_REQUIRE(world >= 2 && world <= KM_XCHG_MAX_WORLD && rank >= 0 && rank < world, ET_ERR_BADARG,
             "et_kmeans_lloyd_sharded: rank %d / world %d outside [2, %d]", rank, world, KM_XCHG_MAX_WORLD);
  ET_REQUIRE(l <= 32, ET_ERR_UNSUPPORTED, "et_kmeans_lloyd_sharded: batch l = %d > 32", l);
  KmLloyd fit{};
  fit.max_iter = max_iter;
  fit.tol = tol;
  fit.cent_out = centroids_out;
  fit.err = err;
  fit.status = status;
  fit.simsum_last = simsum_last;
  fit.labels_final = labels;
  fit.rank = rank;
  fit.world = world;
  fit.xchg = reinterpret_cast<unsigned char* const*>(exchange_peers);
  fit.stamp_base = stamp_base;
  if (n_global > 0) {
    ET_REQUIRE(row_offset >= 0 && n_global >= row_offset + n_local && km_check(l, d, n_global, k_clusters) == ET_OK, ET_ERR_BADARG,
               "et_kmeans_lloyd_sharded: columns [%lld, %lld) outside a data set of %lld", (long long)row_offset,
               (long long)(row_offset + n_local), (long long)n_global);
    fit.col_offset = row_offset;
    fit.n_global = n_global;
  }
  // an empty shard still takes part in every exchange: give the kernel a valid (never dereferenced) data pointer
  return km_dispatch(data ? data : centroids, centroids, l, d, n_local, k_clusters, nullptr, nullptr, nullptr, nullptr, nullptr,
                     workspace, nullptr, nullptr, fit, as_stream(stream));
}

int et_kmeans_accumulate(const float* data, const int64_t* labels, int l, int d, int64_t n, int k_clusters,
                         double* sums, double* counts, void* workspace, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  if (n == 0) return ET_OK;
  ET_REQUIRE((data && labels) || n == 0, ET_ERR_BADARG, "et_kmeans_accumulate: data / labels null");
  ET_REQUIRE(sums && counts && workspace, ET_ERR_BADARG, "et_kmeans_accumulate: sums / counts / workspace null");
  cudaStream_t st = as_stream(stream);
  return km_dispatch(data, nullptr, l, d, n, k_clusters, nullptr, nullptr, sums, counts, nullptr, workspace, nullptr, labels,
                     KmLloyd{}, st);
}

int et_kmeans_finalize(double* sums, double* counts, int l, int d, int k_clusters, const float* old_centroids,
                       float* new_centroids, double* err, double tol, int32_t* status, double* simsum,
                       double* simsum_last, et_stream_t stream) {
  int rc = km_check(l, d, 0, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(sums && counts && new_centroids, ET_ERR_BADARG, "et_kmeans_finalize: null pointer");
  kmeans_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(sums, counts, l, d, k_clusters, old_centroids, new_centroids,
                                                          err, tol, status, simsum, simsum_last);
  return check_launch("kmeans_finalize_kernel");
}

int et_kmeans_farthest_init(const float* data, int l, int d, int64_t n, int k_clusters, int64_t first_index,
                            float* centroids, unsigned long long* scratch, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(data && centroids && scratch, ET_ERR_BADARG, "et_kmeans_farthest_init: null pointer");
  ET_REQUIRE(n >= 1 && first_index >= 0 && first_index < n, ET_ERR_BADARG,
             "et_kmeans_farthest_init: first_index %lld outside [0, N = %lld)", (long long)first_index, (long long)n);
  cudaStream_t st = as_stream(stream);
  // all K - 1 steps in one persistent launch while the points fit the SMs' shared memory (<= ~1.3e6 six-dimensional
  // points on 148 SMs); otherwise one launch per step
  rc = seed_persistent(data, l, d, n, k_clusters, first_index, scratch, centroids, SeedShard{}, st);
  if (rc <= 0) return rc;
  kmeans_seed_init_kernel<<<(l * k_clusters + SEED_SCRATCH_EXTRA + 255) / 256, 256, 0, st>>>(scratch, l, k_clusters, first_index);
  if ((rc = check_launch("kmeans_seed_init_kernel"))) return rc;
  dim3 grid(km_grid(n), l);
  for (int i = 1; i < k_clusters; ++i) {
    launch_seed_step(grid, st, data, d, n, k_clusters, i, scratch, nullptr, nullptr);
    if ((rc = check_launch("kmeans_seed_step_kernel"))) return rc;
  }
  kmeans_seed_gather_kernel<<<(l * d * k_clusters + 255) / 256, 256, 0, st>>>(data, l, d, n, k_clusters, scratch, centroids);
  return check_launch("kmeans_seed_gather_kernel");
}

size_t et_kmeans_d2_workspace_bytes(int l, int trials) {
  if (l < 1 || trials < 1 || trials > SEEDD_TRIALS_MAX) return 0;
  // 128 B of barrier counters, l * trials candidate indices, l * (co-resident blocks) * trials partial sums
  return 128 + (size_t)l * trials * sizeof(long long) + (size_t)l * sm_count() * (trials + 1) * sizeof(double);
}

int et_kmeans_d2_init(const float* data, int l, int d, int64_t n, int k_clusters, int trials, const double* uniform,
                      float* centroids, void* workspace, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(data && uniform && centroids && workspace, ET_ERR_BADARG, "et_kmeans_d2_init: null pointer");
  ET_REQUIRE(n >= 1, ET_ERR_BADARG, "et_kmeans_d2_init: no points");
  ET_REQUIRE(trials >= 1 && trials <= SEEDD_TRIALS_MAX, ET_ERR_BADARG, "et_kmeans_d2_init: trials = %d outside [1, %d]", trials,
             SEEDD_TRIALS_MAX);
  cudaStream_t st = as_stream(stream);
  if (d == 6 && k_clusters <= 32) return seed_d2_launch<6, 32, true>(data, l, d, n, k_clusters, trials, uniform, centroids, workspace, st);
  if (d <= 8 && k_clusters <= 32) return seed_d2_launch<8, 32, false>(data, l, d, n, k_clusters, trials, uniform, centroids, workspace, st);
  return seed_d2_launch<ET_MAX_KM_DIM, ET_MAX_CLUSTERS, false>(data, l, d, n, k_clusters, trials, uniform, centroids, workspace, st);
}

int et_kmeans_farthest_init_sharded(const float* data, int l, int d, int64_t n_local, int k_clusters, int64_t first_global_index,
                                    int64_t row_offset, int64_t n_global, float* centroids, unsigned long long* scratch, int rank,
                                    int world, void* const* exchange_peers, unsigned stamp_base, et_stream_t stream) {
  int rc = km_check(l, d, n_local, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(centroids && scratch && exchange_peers && (n_local == 0 || data), ET_ERR_BADARG,
             "et_kmeans_farthest_init_sharded: null pointer");
  ET_REQUIRE(world >= 2 && world <= KM_XCHG_MAX_WORLD && rank >= 0 && rank < world, ET_ERR_BADARG,
             "et_kmeans_farthest_init_sharded: rank %d / world %d outside [2, %d]", rank, world, KM_XCHG_MAX_WORLD);
  ET_REQUIRE(l <= 32, ET_ERR_UNSUPPORTED, "et_kmeans_farthest_init_sharded: batch l = %d > 32", l);
  ET_REQUIRE(row_offset >= 0 && n_global >= row_offset + n_local && km_check(l, d, n_global, k_clusters) == ET_OK &&
                 first_global_index >= 0 && first_global_index < n_global,
             ET_ERR_BADARG, "et_kmeans_farthest_init_sharded: columns [%lld, %lld), first index %lld outside a data set of %lld",
             (long long)row_offset, (long long)(row_offset + n_local), (long long)first_global_index, (long long)n_global);
  SeedShard sh{};
  sh.rank = rank;
  sh.world = world;
  sh.xchg = reinterpret_cast<unsigned char* const*>(exchange_peers);
  sh.stamp_base = stamp_base;
  sh.col_offset = row_offset;
  sh.n_global = n_global;
  rc = seed_persistent(data ? data : centroids, l, d, n_local, k_clusters, first_global_index, scratch, centroids, sh,
                       as_stream(stream));
  if (rc > 0) return fail(ET_ERR_UNSUPPORTED, "et_kmeans_farthest_init_sharded: %lld local points do not fit the resident kernel",
                          (long long)n_local);
  return rc;
}

static int seed_step_impl(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters, int ncols,
                          unsigned long long* key_out, int64_t row_offset, int64_t n_global, et_stream_t stream);

int et_kmeans_seed_candidate(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters, int ncols,
                             int64_t row_offset, int64_t n_global, long long* gkey_out, et_stream_t stream) {
  ET_REQUIRE(row_offset >= 0 && row_offset + n <= ((int64_t)1 << 32) && (n_global == 0 || n_global >= row_offset + n),
             ET_ERR_UNSUPPORTED, "et_kmeans_seed_candidate: columns [%lld, %lld) of %lld: global indices must stay below 2^32",
             (long long)row_offset, (long long)(row_offset + n), (long long)n_global);
  int rc = seed_step_impl(data, centroids, l, d, n, k_clusters, ncols, reinterpret_cast<unsigned long long*>(gkey_out), row_offset,
                          n_global, stream);
  if (rc) return rc;
  kmeans_seed_globalize_kernel<<<(l + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(gkey_out), l,
                                                                            (unsigned long long)row_offset);
  return check_launch("kmeans_seed_globalize_kernel");
}

int et_kmeans_seed_fetch(const float* data, int l, int d, int64_t n, int64_t row_offset, const long long* gkey,
                         double* coords, et_stream_t stream) {
  int rc = km_check(l, d, n, 1);
  if (rc) return rc;
  ET_REQUIRE(gkey && coords && (n == 0 || data), ET_ERR_BADARG, "et_kmeans_seed_fetch: null pointer");
  kmeans_seed_fetch_kernel<<<(l * d + 255) / 256, 256, 0, as_stream(stream)>>>(data, l, d, n, row_offset, gkey, coords);
  return check_launch("kmeans_seed_fetch_kernel");
}

int et_kmeans_seed_step(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters, int ncols,
                        unsigned long long* key_out, et_stream_t stream) {
  return seed_step_impl(data, centroids, l, d, n, k_clusters, ncols, key_out, 0, 0, stream);
}

static int seed_step_impl(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters, int ncols,
                          unsigned long long* key_out, int64_t row_offset, int64_t n_global, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(key_out && (n == 0 || (data && centroids)), ET_ERR_BADARG, "et_kmeans_seed_step: null pointer");
  ET_REQUIRE(ncols >= 1 && ncols < k_clusters, ET_ERR_BADARG, "et_kmeans_seed_step: ncols = %d outside [1, K)", ncols);
  cudaStream_t st = as_stream(stream);
  fill_u64_kernel<<<(l + 255) / 256, 256, 0, st>>>(key_out, l, ~0ull);   // an empty shard proposes "no candidate"
  if ((rc = check_launch("fill_u64_kernel"))) return rc;
  if (n == 0) return ET_OK;
  dim3 grid(km_grid(n), l);
  launch_seed_step(grid, st, data, d, n, k_clusters, ncols, nullptr, centroids, key_out, row_offset, n_global);
  return check_launch("kmeans_seed_step_kernel");
}

}  // extern "C"
