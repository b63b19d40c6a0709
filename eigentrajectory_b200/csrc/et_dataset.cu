// Dataset preprocessing, host side (reference: utils/dataloader.py:121-232, read_file / poly_fit /
// TrajectoryDataset.__init__).  The reference spends seconds per split in Python loops over numpy masks; this is the
// same windowing as two linear passes in C++ (milliseconds), feeding the (N, T, 2) float32 tensors the descriptor
// kernels consume.  Nothing here touches the GPU: every pointer is a HOST pointer.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "et_common.cuh"

namespace et {

static inline bool py_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// numpy.around(x, 4) for float64: rint(x * 1e4) / 1e4 (numpy multiplies, rounds half to even, divides)
static inline double around4(double v) { return nearbyint(v * 10000.0) / 10000.0; }

// Sum of squared residuals of the least-squares quadratic through (t = 0..n-1, y): what
// numpy.polyfit(t, y, 2, full=True)[1] returns (dataloader.py:144-147).  Projection on an orthonormal basis of
// {1, t, t^2} built by modified Gram-Schmidt in long double; y is shifted by y[0] first (the constant is in the span).
static double quadratic_residual(const double* y, int n) {
  if (n <= 3) return 0.0;
  std::vector<long double> q0(n), q1(n), q2(n), r(n);
  const long double mid = 0.5L * (n - 1);
  for (int i = 0; i < n; ++i) { q0[i] = 1.0L; q1[i] = i - mid; q2[i] = (i - mid) * (i - mid); }
  auto dot = [&](const std::vector<long double>& a, const std::vector<long double>& b) {
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  };
  auto normalise = [&](std::vector<long double>& a) {
    const long double nr = sqrtl(dot(a, a));
    for (int i = 0; i < n; ++i) a[i] /= nr;
  };
  auto remove = [&](std::vector<long double>& a, const std::vector<long double>& q) {
    const long double c = dot(a, q);
    for (int i = 0; i < n; ++i) a[i] -= c * q[i];
  };
  normalise(q0);
  remove(q1, q0); normalise(q1);
  remove(q2, q0); remove(q2, q1); normalise(q2);
  for (int i = 0; i < n; ++i) r[i] = (long double)y[i] - (long double)y[0];
  remove(r, q0); remove(r, q1); remove(r, q2);
  remove(r, q0); remove(r, q1); remove(r, q2);   // second pass: re-orthogonalisation
  return (double)dot(r, r);
}

}  // namespace et

using namespace et;

extern "C" {

int et_dataset_parse_host(const char* text_host, size_t len, char delim, double* rows_host, int64_t capacity,
                          int64_t* n_rows_host) {
  ET_REQUIRE(n_rows_host, ET_ERR_BADARG, "et_dataset_parse_host: n_rows_host is null");
  ET_REQUIRE(text_host || len == 0, ET_ERR_BADARG, "et_dataset_parse_host: text is null");
  int64_t n = 0;
  size_t pos = 0;
  int64_t line_no = 0;
  while (pos < len) {
    size_t end = pos;
    while (end < len && text_host[end] != '\n') ++end;
    ++line_no;
    // line.strip()
    size_t a = pos, b = end;
    while (a < b && py_space(text_host[a])) ++a;
    while (b > a && py_space(text_host[b - 1])) --b;
    // .split(delim) -> float(token) for every token; exactly four columns (frame, ped, x, y)
    double v[4];
    int col = 0;
    size_t t0 = a;
    for (size_t i = a; i <= b; ++i) {
      if (i == b || text_host[i] == delim) {
        size_t s = t0, e = i;
        while (s < e && py_space(text_host[s])) ++s;
        while (e > s && py_space(text_host[e - 1])) --e;
        char buf[64];
        if (e == s || e - s >= sizeof(buf))
          return fail(ET_ERR_BADARG, "et_dataset_parse_host: line %lld: empty or oversized field", (long long)line_no);
        if (col >= 4)
          return fail(ET_ERR_BADARG, "et_dataset_parse_host: line %lld: more than 4 columns", (long long)line_no);
        memcpy(buf, text_host + s, e - s);
        buf[e - s] = 0;
        char* stop = nullptr;
        v[col] = strtod(buf, &stop);
        if (stop != buf + (e - s))
          return fail(ET_ERR_BADARG, "et_dataset_parse_host: line %lld: could not convert '%s' to float", (long long)line_no, buf);
        ++col;
        t0 = i + 1;
      }
    }
    if (col != 4)
      return fail(ET_ERR_BADARG, "et_dataset_parse_host: line %lld: %d columns, expected 4 (frame, ped, x, y)",
                  (long long)line_no, col);
    if (rows_host) {
      if (n >= capacity) return fail(ET_ERR_BADARG, "et_dataset_parse_host: more than capacity = %lld rows", (long long)capacity);
      memcpy(rows_host + 4 * n, v, sizeof(v));
    }
    ++n;
    pos = end + 1;
  }
  *n_rows_host = n;
  return ET_OK;
}

int et_dataset_windows_host(const double* rows_host, int64_t n_rows, int obs_len, int pred_len, int skip, double threshold,
                            int min_ped, float* traj_host, float* non_linear_host, int32_t* peds_in_seq_host,
                            int64_t cap_peds, int64_t cap_seq, int64_t* n_peds_host, int64_t* n_seq_host) {
  ET_REQUIRE(rows_host && n_peds_host && n_seq_host, ET_ERR_BADARG, "et_dataset_windows_host: null pointer");
  ET_REQUIRE(n_rows >= 1, ET_ERR_BADARG, "et_dataset_windows_host: no rows (the reference fails on an empty file)");
  ET_REQUIRE(obs_len >= 1 && pred_len >= 1 && skip >= 1, ET_ERR_BADARG, "et_dataset_windows_host: obs_len / pred_len / skip < 1");
  const int L = obs_len + pred_len;
  const bool fill = traj_host != nullptr;
  ET_REQUIRE(!fill || (non_linear_host && peds_in_seq_host), ET_ERR_BADARG,
             "et_dataset_windows_host: traj given without non_linear / peds_in_seq");

  // frames = np.unique(data[:, 0]); frame_data[f] = rows of frame f in file order  (dataloader.py:189-192)
  std::vector<double> frames(n_rows);
  for (int64_t i = 0; i < n_rows; ++i) {
    const double f = rows_host[4 * i];
    ET_REQUIRE(f == f && rows_host[4 * i + 1] == rows_host[4 * i + 1], ET_ERR_BADARG,
               "et_dataset_windows_host: row %lld: NaN frame or pedestrian id", (long long)i);
    // the reference looks frames up by their 4-decimal rounding (frames.index(around(frame, 4)), dataloader.py:206-207)
    ET_REQUIRE(around4(f) == f, ET_ERR_BADARG,
               "et_dataset_windows_host: row %lld: frame id %.17g changes under 4-decimal rounding", (long long)i, f);
    frames[i] = f;
  }
  std::sort(frames.begin(), frames.end());
  frames.erase(std::unique(frames.begin(), frames.end()), frames.end());
  const int64_t F = (int64_t)frames.size();
  std::vector<int64_t> fidx(n_rows), start(F + 1, 0);
  for (int64_t i = 0; i < n_rows; ++i) {
    fidx[i] = std::lower_bound(frames.begin(), frames.end(), rows_host[4 * i]) - frames.begin();
    ++start[fidx[i] + 1];
  }
  for (int64_t f = 0; f < F; ++f) start[f + 1] += start[f];
  std::vector<int64_t> by_frame(n_rows), cursor(start.begin(), start.end() - 1);
  for (int64_t i = 0; i < n_rows; ++i) by_frame[cursor[fidx[i]]++] = i;

  // num_sequences = int(math.ceil((len(frames) - seq_len + 1) / skip)); idx in range(0, num_sequences * skip + 1, skip)
  const int64_t num_sequences = (int64_t)ceil((double)(F - L + 1) / (double)skip);
  const int64_t idx_end = num_sequences * skip + 1;

  int64_t n_peds = 0, n_seq = 0;
  struct Entry { double ped; int64_t pos; };   // pos = position inside the window (frame-major, file order)
  std::vector<Entry> win;
  std::vector<double> xs(L), ys(L);
  std::vector<float> seq_traj, seq_nl;
  for (int64_t idx = 0; idx < idx_end; idx += skip) {
    ET_REQUIRE(idx < F, ET_ERR_BADARG,
               "et_dataset_windows_host: window %lld starts past the last frame (np.concatenate of nothing in the reference)",
               (long long)idx);
    const int64_t f_hi = std::min<int64_t>(idx + L, F);
    const int64_t lo = start[idx], hi = start[f_hi];
    win.resize(hi - lo);
    for (int64_t p = lo; p < hi; ++p) win[p - lo] = Entry{rows_host[4 * by_frame[p] + 1], p};
    // peds_in_curr_seq = np.unique(...): ascending ids; rows of one pedestrian keep window order
    std::stable_sort(win.begin(), win.end(), [](const Entry& a, const Entry& b) { return a.ped < b.ped; });
    seq_traj.clear();
    seq_nl.clear();
    int considered = 0;
    for (size_t g0 = 0; g0 < win.size();) {
      size_t g1 = g0;
      while (g1 < win.size() && win[g1].ped == win[g0].ped) ++g1;
      const int64_t first = fidx[by_frame[win[g0].pos]], last = fidx[by_frame[win[g1 - 1].pos]];
      const int64_t pad_front = first - idx, pad_end = last - idx + 1;
      if (pad_end - pad_front == L) {
        // curr_seq[_idx, :, pad_front:pad_end] = curr_ped_seq needs exactly seq_len rows (numpy raises otherwise)
        ET_REQUIRE((int64_t)(g1 - g0) == L, ET_ERR_BADARG,
                   "et_dataset_windows_host: pedestrian %.10g spans window %lld but has %lld rows instead of %d "
                   "(missing or duplicated frames; the reference raises ValueError)",
                   win[g0].ped, (long long)idx, (long long)(g1 - g0), L);
        for (int t = 0; t < L; ++t) {
          const double* r = rows_host + 4 * by_frame[win[g0 + t].pos];
          xs[t] = around4(r[2]);     // np.around(curr_ped_seq, decimals=4)
          ys[t] = around4(r[3]);
        }
        if (fill) {
          for (int t = 0; t < L; ++t) { seq_traj.push_back((float)xs[t]); seq_traj.push_back((float)ys[t]); }
          // poly_fit(curr_ped_seq, pred_len, threshold): quadratic fit of the last pred_len frames (dataloader.py:135-151)
          const double res = quadratic_residual(xs.data() + L - pred_len, pred_len) +
                             quadratic_residual(ys.data() + L - pred_len, pred_len);
          seq_nl.push_back((pred_len > 3 && res >= threshold) ? 1.0f : 0.0f);
        }
        ++considered;
      }
      g0 = g1;
    }
    if (considered > min_ped) {     // strictly greater, as dataloader.py:221
      if (fill) {
        ET_REQUIRE(n_peds + considered <= cap_peds && n_seq + 1 <= cap_seq, ET_ERR_BADARG,
                   "et_dataset_windows_host: output capacity exceeded (%lld pedestrians / %lld sequences)",
                   (long long)cap_peds, (long long)cap_seq);
        memcpy(traj_host + n_peds * L * 2, seq_traj.data(), seq_traj.size() * sizeof(float));
        memcpy(non_linear_host + n_peds, seq_nl.data(), seq_nl.size() * sizeof(float));
        peds_in_seq_host[n_seq] = considered;
      }
      n_peds += considered;
      n_seq += 1;
    }
  }
  *n_peds_host = n_peds;
  *n_seq_host = n_seq;
  return ET_OK;
}

}  // extern "C"
