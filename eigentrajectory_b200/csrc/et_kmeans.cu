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
// The update half accumulates, per warp, at most 32 points in fp32 shared-memory atomics and folds them
// into float64 registers after every batch of 32, so centroid sums carry fp64 accuracy.
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

// best similarity / label of one point against ncols staged centroids (cs: [d][kpitch], bn: [ncols])
template <int DMAX>
__device__ __forceinline__ void best_centroid(const float (&a)[DMAX], int d, float anorm, const float* cs, int kpitch,
                                              const float* bn, int ncols, float& best, int& label) {
  best = 0.f;
  label = 0;
  for (int j = 0; j < ncols; ++j) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < DMAX; ++i)
      if (i < d) dot = fmaf(a[i], cs[i * kpitch + j], dot);
    const float y = __fsub_rn(__fsub_rn(__fmul_rn(dot, 2.0f), anorm), bn[j]);
    if (j == 0 || y > best || (y != y && best == best)) { best = y; label = j; }
  }
}

constexpr int KM_WARPS = 8;
constexpr int KM_THREADS = KM_WARPS * 32;

// workspace: l tickets (uint32, zero on entry / exit) padded to 128 B, then l * gridDim.x partial records of
// (d*K + K + 1) doubles.
template <int DMAX, int KMAX>
__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(
    const float* __restrict__ data, const float* __restrict__ centroids, int d, int64_t n, int k,
    int64_t* __restrict__ labels, float* __restrict__ maxsims, double* __restrict__ sums, double* __restrict__ counts,
    double* __restrict__ simsum, unsigned* __restrict__ tickets, double* __restrict__ partials,
    const int32_t* __restrict__ status, const int64_t* __restrict__ labels_in) {
  if (status && status[0] != 0) return;   // converged on an earlier iteration: nothing to do
  constexpr int REC = KMAX * (DMAX + 1);       // per-warp fp32 record: [cluster][d sums..., count]
  constexpr int NQ = (REC + 31) / 32;
  __shared__ float cs[DMAX * KMAX];
  __shared__ float bn[KMAX];
  __shared__ float wrec[KM_WARPS][REC];
  __shared__ double blk[REC + 1];
  __shared__ unsigned is_last;
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rec = k * (d + 1);
  const bool accumulate = sums != nullptr;

  if (centroids)
    for (int e = tid; e < d * k; e += KM_THREADS) cs[(e / k) * KMAX + (e % k)] = __ldg(centroids + (int64_t)l * d * k + e);
  for (int e = tid; e < REC; e += KM_THREADS) {
#pragma unroll
    for (int w = 0; w < KM_WARPS; ++w) wrec[w][e] = 0.f;
  }
  __syncthreads();
  if (tid < k) {
    float v[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) v[i] = (i < d) ? cs[i * KMAX + tid] : 0.f;
    bn[tid] = sumsq_torch_order<DMAX>(v, d, col_is_sequential(tid, k));
  }
  __syncthreads();

  const float* dl = data + (int64_t)l * d * n;
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
  double sim_acc = 0.0;
  float* mine = wrec[warp];

  const int64_t stride = (int64_t)gridDim.x * KM_THREADS;
  // all lanes of a warp iterate together (the loop bound is warp-uniform)
  for (int64_t base = (int64_t)blockIdx.x * KM_THREADS + warp * 32; base < n; base += stride) {
    const int64_t i = base + lane;
    if (i < n) {
      float a[DMAX];
#pragma unroll
      for (int r = 0; r < DMAX; ++r) a[r] = (r < d) ? __ldg(dl + (int64_t)r * n + i) : 0.f;
      float best = 0.f;
      int label;
      if (labels_in) {   // compute_centroids with caller-supplied labels: accumulation only
        const int64_t li = labels_in[(int64_t)l * n + i];
        label = (li >= 0 && li < k) ? (int)li : -1;
      } else {
        const float anorm = sumsq_torch_order<DMAX>(a, d, col_is_sequential(i, n));
        best_centroid<DMAX>(a, d, anorm, cs, KMAX, bn, k, best, label);
      }
      if (labels) labels[(int64_t)l * n + i] = label;
      if (maxsims) maxsims[(int64_t)l * n + i] = best;
      if (accumulate && label >= 0) {
        float* slot = mine + label * (d + 1);
#pragma unroll
        for (int r = 0; r < DMAX; ++r)
          if (r < d) atomicAdd(slot + r, a[r]);
        atomicAdd(slot + d, 1.0f);
        sim_acc += (double)best;
      }
    }
    if (accumulate) {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = lane + 32 * q;
        if (e < rec) {
          acc[q] += (double)mine[e];
          mine[e] = 0.f;
        }
      }
      __syncwarp();
    }
  }
  if (!accumulate) return;

  // ---- block reduction in fixed warp order ----
  for (int e = tid; e <= REC; e += KM_THREADS) blk[e] = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sim_acc += __shfl_xor_sync(0xffffffffu, sim_acc, o);
  __syncthreads();
  for (int w = 0; w < KM_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int e = lane + 32 * q;
        if (e < rec) blk[e] += acc[q];
      }
      if (lane == 0) blk[REC] += sim_acc;
    }
    __syncthreads();
  }
  const int out_rec = rec + 1;
  double* part = partials + ((size_t)l * gridDim.x + blockIdx.x) * out_rec;
  for (int e = tid; e < out_rec; e += KM_THREADS) part[e] = (e < rec) ? blk[e] : blk[REC];
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&tickets[l], 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- the last block of this batch entry folds the partial records in block order ----
  for (int e = tid; e < out_rec; e += KM_THREADS) {
    double s0 = 0.0, s1 = 0.0;
    unsigned b = 0;
    for (; b + 2 <= gridDim.x; b += 2) {
      s0 += __ldcg(partials + ((size_t)l * gridDim.x + b) * out_rec + e);
      s1 += __ldcg(partials + ((size_t)l * gridDim.x + b + 1) * out_rec + e);
    }
    if (b < gridDim.x) s0 += __ldcg(partials + ((size_t)l * gridDim.x + b) * out_rec + e);
    const double tot = s0 + s1;
    if (e < rec) {
      const int c = e / (d + 1), r = e % (d + 1);
      if (r < d) sums[((int64_t)l * d + r) * k + c] += tot;
      else counts[(int64_t)l * k + c] += tot;
    } else if (simsum) {
      simsum[l] += tot;
    }
  }
  if (tid == 0) tickets[l] = 0u;
}

// new = float(sums / counts); err = sum (old - new)^2; clears the accumulators for the next iteration.
__global__ void kmeans_finalize_kernel(double* __restrict__ sums, double* __restrict__ counts, int l, int d, int k,
                                       const float* __restrict__ old_c, float* __restrict__ new_c,
                                       double* __restrict__ err, double tol, int32_t* __restrict__ status) {
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

__global__ void kmeans_seed_init_kernel(unsigned long long* scratch, int l, int k, int64_t first_index) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < l * k) scratch[e] = (e % k == 0) ? (unsigned long long)first_index : ~0ull;
}

// Step `ncols` (1 <= ncols < K): the chosen points scratch[l*K + 0 .. ncols) are the current centroids; every
// point takes its best similarity against them exactly as the reference recomputes it, and the point with the
// lowest best similarity (lowest index on ties) is recorded in scratch[l*K + ncols].
template <int DMAX, int KMAX>
__global__ void __launch_bounds__(KM_THREADS) kmeans_seed_step_kernel(const float* __restrict__ data, int d, int64_t n,
                                                                      int k, int ncols,
                                                                      unsigned long long* __restrict__ scratch) {
  __shared__ float cs[DMAX * KMAX];
  __shared__ float bn[KMAX];
  __shared__ unsigned long long wmin[KM_WARPS];
  const int l = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* dl = data + (int64_t)l * d * n;
  for (int e = tid; e < d * ncols; e += KM_THREADS) {
    const int r = e / ncols, j = e % ncols;
    const int64_t idx = (int64_t)(scratch[(int64_t)l * k + j] & 0xffffffffull);
    cs[r * KMAX + j] = __ldg(dl + (int64_t)r * n + idx);
  }
  __syncthreads();
  if (tid < ncols) {
    float v[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) v[i] = (i < d) ? cs[i * KMAX + tid] : 0.f;
    bn[tid] = sumsq_torch_order<DMAX>(v, d, col_is_sequential(tid, ncols));
  }
  __syncthreads();
  unsigned long long key = ~0ull;
  for (int64_t i = (int64_t)blockIdx.x * KM_THREADS + tid; i < n; i += (int64_t)gridDim.x * KM_THREADS) {
    float a[DMAX];
#pragma unroll
    for (int r = 0; r < DMAX; ++r) a[r] = (r < d) ? __ldg(dl + (int64_t)r * n + i) : 0.f;
    const float anorm = sumsq_torch_order<DMAX>(a, d, col_is_sequential(i, n));
    float best;
    int label;
    best_centroid<DMAX>(a, d, anorm, cs, KMAX, bn, ncols, best, label);
    const unsigned long long kk = pack_min_key(best, i);
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
    atomicMin(&scratch[(int64_t)l * k + ncols], key);
  }
}

__global__ void kmeans_seed_gather_kernel(const float* __restrict__ data, int l, int d, int64_t n, int k,
                                          const unsigned long long* __restrict__ scratch, float* __restrict__ centroids) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= l * d * k) return;
  const int j = e % k, r = (e / k) % d, li = e / (d * k);
  const int64_t idx = (int64_t)(scratch[(int64_t)li * k + j] & 0xffffffffull);
  centroids[e] = __ldg(data + ((int64_t)li * d + r) * n + idx);
}

static int km_grid(int64_t n) {
  int64_t g = (n + KM_THREADS - 1) / KM_THREADS;
  const int64_t cap = (int64_t)sm_count() * 4;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

static int km_check(int l, int d, int64_t n, int k) {
  if (l < 1 || l > 65535) return fail(ET_ERR_UNSUPPORTED, "k-means: batch l = %d outside [1, 65535]", l);
  if (d < 1 || d > ET_MAX_KM_DIM) return fail(ET_ERR_UNSUPPORTED, "k-means: d = %d outside [1, %d]", d, ET_MAX_KM_DIM);
  if (k < 1 || k > ET_MAX_CLUSTERS) return fail(ET_ERR_UNSUPPORTED, "k-means: K = %d outside [1, %d]", k, ET_MAX_CLUSTERS);
  if (n < 0 || n >= ((int64_t)1 << 32)) return fail(ET_ERR_UNSUPPORTED, "k-means: N = %lld outside [0, 2^32)", (long long)n);
  return ET_OK;
}

}  // namespace et

using namespace et;

extern "C" {

size_t et_kmeans_workspace_bytes(int l, int d, int k_clusters) {
  if (l < 1 || d < 1 || k_clusters < 1) return 0;
  const size_t tickets = (((size_t)l * 4 + 127) / 128) * 128;
  return tickets + (size_t)l * sm_count() * 4 * ((size_t)k_clusters * (d + 1) + 1) * sizeof(double);
}

int et_kmeans_assign(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters,
                     int64_t* labels, float* maxsims, double* sums, double* counts, double* simsum, void* workspace,
                     const int32_t* status, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE((data && centroids) || n == 0, ET_ERR_BADARG, "et_kmeans_assign: data / centroids null");
  ET_REQUIRE(!sums || (counts && workspace), ET_ERR_BADARG, "et_kmeans_assign: sums given without counts / workspace");
  if (n == 0) return ET_OK;
  const size_t tickets = (((size_t)l * 4 + 127) / 128) * 128;
  unsigned* tk = reinterpret_cast<unsigned*>(workspace);
  double* parts = workspace ? reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + tickets) : nullptr;
  dim3 grid(km_grid(n), l);
  cudaStream_t st = as_stream(stream);
  if (d <= 8 && k_clusters <= 32)
    kmeans_assign_kernel<8, 32><<<grid, KM_THREADS, 0, st>>>(data, centroids, d, n, k_clusters, labels, maxsims, sums,
                                                           counts, simsum, tk, parts, status, nullptr);
  else
    kmeans_assign_kernel<ET_MAX_KM_DIM, ET_MAX_CLUSTERS><<<grid, KM_THREADS, 0, st>>>(
        data, centroids, d, n, k_clusters, labels, maxsims, sums, counts, simsum, tk, parts, status, nullptr);
  return check_launch("kmeans_assign_kernel");
}

int et_kmeans_accumulate(const float* data, const int64_t* labels, int l, int d, int64_t n, int k_clusters,
                         double* sums, double* counts, void* workspace, et_stream_t stream) {
  int rc = km_check(l, d, n, k_clusters);
  if (rc) return rc;
  ET_REQUIRE((data && labels) || n == 0, ET_ERR_BADARG, "et_kmeans_accumulate: data / labels null");
  ET_REQUIRE(sums && counts && workspace, ET_ERR_BADARG, "et_kmeans_accumulate: sums / counts / workspace null");
  if (n == 0) return ET_OK;
  const size_t tickets = (((size_t)l * 4 + 127) / 128) * 128;
  unsigned* tk = reinterpret_cast<unsigned*>(workspace);
  double* parts = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + tickets);
  dim3 grid(km_grid(n), l);
  cudaStream_t st = as_stream(stream);
  if (d <= 8 && k_clusters <= 32)
    kmeans_assign_kernel<8, 32><<<grid, KM_THREADS, 0, st>>>(data, nullptr, d, n, k_clusters, nullptr, nullptr, sums,
                                                           counts, nullptr, tk, parts, nullptr, labels);
  else
    kmeans_assign_kernel<ET_MAX_KM_DIM, ET_MAX_CLUSTERS><<<grid, KM_THREADS, 0, st>>>(
        data, nullptr, d, n, k_clusters, nullptr, nullptr, sums, counts, nullptr, tk, parts, nullptr, labels);
  return check_launch("kmeans_assign_kernel(accumulate)");
}

int et_kmeans_finalize(double* sums, double* counts, int l, int d, int k_clusters, const float* old_centroids,
                       float* new_centroids, double* err, double tol, int32_t* status, et_stream_t stream) {
  int rc = km_check(l, d, 0, k_clusters);
  if (rc) return rc;
  ET_REQUIRE(sums && counts && new_centroids, ET_ERR_BADARG, "et_kmeans_finalize: null pointer");
  kmeans_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(sums, counts, l, d, k_clusters, old_centroids, new_centroids,
                                                          err, tol, status);
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
  kmeans_seed_init_kernel<<<(l * k_clusters + 255) / 256, 256, 0, st>>>(scratch, l, k_clusters, first_index);
  if ((rc = check_launch("kmeans_seed_init_kernel"))) return rc;
  dim3 grid(km_grid(n), l);
  for (int i = 1; i < k_clusters; ++i) {
    if (d <= 8 && k_clusters <= 32)
      kmeans_seed_step_kernel<8, 32><<<grid, KM_THREADS, 0, st>>>(data, d, n, k_clusters, i, scratch);
    else
      kmeans_seed_step_kernel<ET_MAX_KM_DIM, ET_MAX_CLUSTERS><<<grid, KM_THREADS, 0, st>>>(data, d, n, k_clusters, i, scratch);
    if ((rc = check_launch("kmeans_seed_step_kernel"))) return rc;
  }
  kmeans_seed_gather_kernel<<<(l * d * k_clusters + 255) / 256, 256, 0, st>>>(data, l, d, n, k_clusters, scratch, centroids);
  return check_launch("kmeans_seed_gather_kernel");
}

}  // extern "C"
