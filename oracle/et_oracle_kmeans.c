/*
 * Plain-C restatement of BatchKMeans.get_labels / euc_sim (reference: EigenTrajectory/kmeans.py:59-76,143-158) as the
 * reference's CPU arithmetic evaluates it (torch 2.11, MKL) -- TEST INFRASTRUCTURE ONLY, the checker for the CUDA
 * kernel's bit-exact claim.  Built by `make oracle` / __graft_entry__.build() into oracle/_build/libet_oracle.so and
 * pinned in tests/test_oracle_golden.py against the lock-step trace the unmodified reference produced
 * (tests/golden/kmeans.npz: labels and similarities bit for bit).  Nothing in the product links or loads it.
 *
 *   sim[n][j] = fl(fl(fl(2 * dot) - |a_n|^2) - |b_j|^2)           (in-place mul_, sub_, sub_: kmeans.py:71-74)
 *   dot       = one ascending fp32 FMA chain from 0                (what MKL sgemm does for d <= 16)
 *   |v|^2     = separately rounded squares, summed the way ATen's sum(dim=-2) does for that column: sequentially for
 *               columns inside a full block of 32 columns, with four interleaved partial sums ((p0+p1)+p2)+p3 (p0 also
 *               takes the rows beyond the last full group of four) for tail columns; a tensor of 4..7 columns sums
 *               its first four columns sequentially
 *   label     = arg-max over j, lowest index on ties, NaN wins      (torch.max, kmeans.py:156)
 *
 * Compile without contraction and without fast-math: gcc -O2 -ffp-contract=off -shared -fPIC.
 */
#include <math.h>
#include <stdint.h>

static int col_is_sequential(int64_t idx, int64_t ncols) {
  if (ncols >= 4 && ncols < 8) return idx < 4;
  return idx < 32 * (ncols / 32);
}

/* sum of squares of v[0..d) with stride `stride` between rows, in torch's order for this column */
static float sumsq_torch_order(const float* v, int64_t stride, int d, int sequential) {
  if (sequential) {
    float acc = 0.0f;
    for (int i = 0; i < d; ++i) {
      const float sq = v[i * stride] * v[i * stride];
      acc = acc + sq;
    }
    return acc;
  }
  float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int full = d / 4;
  for (int g = 0; g < full; ++g)
    for (int q = 0; q < 4; ++q) {
      const float x = v[(4 * g + q) * stride];
      const float sq = x * x;
      p[q] = p[q] + sq;
    }
  for (int i = 4 * full; i < d; ++i) {
    const float sq = v[i * stride] * v[i * stride];
    p[0] = p[0] + sq;
  }
  return ((p[0] + p[1]) + p[2]) + p[3];
}

/* data (l, d, n), centroids (l, d, k) contiguous; labels (l, n) int64, maxsims (l, n) float */
void et_oracle_kmeans_assign(const float* data, const float* centroids, int l, int d, int64_t n, int k, int64_t* labels,
                             float* maxsims) {
  for (int li = 0; li < l; ++li) {
    const float* a = data + (int64_t)li * d * n;
    const float* b = centroids + (int64_t)li * d * k;
    for (int64_t i = 0; i < n; ++i) {
      const float anorm = sumsq_torch_order(a + i, n, d, col_is_sequential(i, n));
      float best = 0.0f;
      int64_t arg = 0;
      for (int j = 0; j < k; ++j) {
        float dot = 0.0f;
        for (int r = 0; r < d; ++r) dot = fmaf(a[(int64_t)r * n + i], b[(int64_t)r * k + j], dot);
        const float bnorm = sumsq_torch_order(b + j, k, d, col_is_sequential(j, k));
        float y = dot * 2.0f;
        y = y - anorm;
        y = y - bnorm;
        /* torch.max: the first NaN wins; otherwise strictly greater replaces (lowest index on ties) */
        if (j == 0 || (y > best) || (isnan(y) && !isnan(best))) {
          best = y;
          arg = j;
        }
      }
      labels[(int64_t)li * n + i] = arg;
      maxsims[(int64_t)li * n + i] = best;
    }
  }
}
