// Mask-free glue of EigenTrajectory.forward (reference: EigenTrajectory/model.py:73-105).
//
// The reference splits every batch into a moving and a static group with boolean-mask gathers (each one a
// `nonzero` + device->host sync), runs the two descriptors / anchors separately and scatters the results back:
// ~290 kernel launches and 10 syncs for a scene of a few dozen pedestrians.  Here the group is decided per
// pedestrian inside the kernel (moving iff ||(last - third_last) / 2|| > static_dist, model.py:46,73) and selects
// the bases, the anchor set and whether the scale stage applies, so projection and anchor+reconstruction are one
// launch each, with no gather/scatter and no host synchronisation.  Per-scene batches are tiny (2..57 pedestrians,
// SURVEY section 0.4): these kernels are written for launch count, not bandwidth -- thread per pedestrian /
// per (pedestrian, sample), any T <= 32, k <= 32.
#include "et_common.cuh"

namespace et {

constexpr int FMAX2T = 2 * ET_MAX_T;

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int count) {
  for (int e = threadIdx.x; e < count; e += blockDim.x) dst[e] = __ldg(src + e);
}

__global__ void forward_project_kernel(const float* __restrict__ obs, const float* __restrict__ pred, int64_t n, int to2,
                                       int tp2, const float* __restrict__ U_obs_m, const float* __restrict__ U_obs_s,
                                       const float* __restrict__ U_pred_m, const float* __restrict__ U_pred_s, int k,
                                       float static_dist, float* __restrict__ C_obs, float* __restrict__ C_pred,
                                       float* __restrict__ ori, float* __restrict__ rot, float* __restrict__ sca,
                                       unsigned char* __restrict__ moving) {
  extern __shared__ float Us[];
  float* Uom = Us;
  float* Uos = Uom + to2 * k;
  float* Upm = Uos + to2 * k;
  float* Ups = Upm + tp2 * k;
  stage(Uom, U_obs_m, to2 * k);
  stage(Uos, U_obs_s, to2 * k);
  if (pred) {
    stage(Upm, U_pred_m, tp2 * k);
    stage(Ups, U_pred_s, tp2 * k);
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[FMAX2T];
  for (int r = 0; r < to2; ++r) x[r] = __ldg(obs + i * to2 + r);
  NormState st = make_norm_state(x[to2 - 2], x[to2 - 1], x[to2 - 6], x[to2 - 5]);
  // model.py:46,73: (obs[-1] - obs[-3]).div(2).norm(p=2) > static_dist
  const float hx = (x[to2 - 2] - x[to2 - 6]) / 2.0f, hy = (x[to2 - 1] - x[to2 - 5]) / 2.0f;
  const bool mv = sqrtf(__fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy))) > static_dist;
  const int flags = ET_NORM_ORI | ET_NORM_ROT | (mv ? ET_NORM_SCA : 0);
  if (!mv) st.sca = 1.0f;   // static rows carry a neutral scale so that reconstruction treats every row alike
  moving[i] = mv ? 1 : 0;
  store_norm_state(ori, rot, sca, i, st, ET_NORM_ORI | ET_NORM_ROT | ET_NORM_SCA);
  for (int pass = 0; pass < 2; ++pass) {
    const int t2 = pass ? tp2 : to2;
    const float* Ub = pass ? (mv ? Upm : Ups) : (mv ? Uom : Uos);
    float* Cout = pass ? C_pred : C_obs;
    if (pass) {
      if (!pred) break;
      for (int r = 0; r < t2; ++r) x[r] = __ldg(pred + i * t2 + r);
    }
    for (int r = 0; r < t2; r += 2) norm_fwd(x[r], x[r + 1], st, flags);
    for (int j = 0; j < k; ++j) {
      float acc = 0.f;
      for (int r = 0; r < t2; ++r) acc = fmaf(Ub[r * k + j], x[r], acc);
      Cout[(int64_t)j * n + i] = acc;
    }
  }
}

// one thread per (pedestrian, sample): out[s,n] = denormalise(U_g (C[:,n,s] + anchor_g[:,s]))
__global__ void forward_reconstruct_kernel(const float* __restrict__ C, const float* __restrict__ anchor_m,
                                           const float* __restrict__ anchor_s, int64_t n, int s, int k, int t2,
                                           const float* __restrict__ U_m, const float* __restrict__ U_s,
                                           const unsigned char* __restrict__ moving, const float* __restrict__ ori,
                                           const float* __restrict__ rot, const float* __restrict__ sca,
                                           float* __restrict__ out) {
  extern __shared__ float Us[];
  float* Um = Us;
  float* Usx = Um + t2 * k;
  stage(Um, U_m, t2 * k);
  stage(Usx, U_s, t2 * k);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * s) return;
  const int64_t i = e / s;
  const int si = (int)(e - i * s);
  const bool mv = moving[i] != 0;
  const float* Ub = mv ? Um : Usx;
  const float* an = mv ? anchor_m : anchor_s;
  const int flags = ET_NORM_ORI | ET_NORM_ROT | ET_NORM_SCA;
  const NormState st = load_norm_state(ori, rot, sca, i, flags);
  const float inv_sca = 1.0f / st.sca;
  float c[ET_MAX_K];
  for (int j = 0; j < k; ++j) {
    float v = __ldg(C + ((int64_t)j * n + i) * s + si);
    if (an) v = __ldg(an + j * s + si) + v;
    c[j] = v;
  }
  float* o = out + ((int64_t)si * n + i) * t2;
  for (int r = 0; r < t2; r += 2) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < k; ++j) {
      a = fmaf(Ub[r * k + j], c[j], a);
      b = fmaf(Ub[(r + 1) * k + j], c[j], b);
    }
    norm_bwd(a, b, st, inv_sca, flags);
    o[r] = a;
    o[r + 1] = b;
  }
}

// grad_C[:,n,s] = U_g^T ((g R) / sca)
__global__ void forward_reconstruct_bwd_kernel(const float* __restrict__ grad_out, int64_t n, int s, int k, int t2,
                                               const float* __restrict__ U_m, const float* __restrict__ U_s,
                                               const unsigned char* __restrict__ moving, const float* __restrict__ rot,
                                               const float* __restrict__ sca, float* __restrict__ grad_C) {
  extern __shared__ float Us[];
  float* Um = Us;
  float* Usx = Um + t2 * k;
  stage(Um, U_m, t2 * k);
  stage(Usx, U_s, t2 * k);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * s) return;
  const int64_t i = e / s;
  const int si = (int)(e - i * s);
  const float* Ub = moving[i] ? Um : Usx;
  const NormState st = load_norm_state(nullptr, rot, sca, i, ET_NORM_ROT | ET_NORM_SCA);
  const float inv_sca = 1.0f / st.sca;
  float g[FMAX2T];
  const float* go = grad_out + ((int64_t)si * n + i) * t2;
  for (int r = 0; r < t2; r += 2) {
    const float a = __ldg(go + r), b = __ldg(go + r + 1);
    g[r] = (a * st.r00 + b * st.r10) * inv_sca;
    g[r + 1] = (a * st.r01 + b * st.r11) * inv_sca;
  }
  for (int j = 0; j < k; ++j) {
    float acc = 0.f;
    for (int r = 0; r < t2; ++r) acc = fmaf(Ub[r * k + j], g[r], acc);
    grad_C[((int64_t)j * n + i) * s + si] = acc;
  }
}

// ---- the three training losses of model.py:119-123 in one launch -----------------------------------------------
//   loss_eigentraj     = mean_n min_s || C_pred[:,n,s] - C_gt[:,n] ||_2          (C_pred = anchor_g + C_refine)
//   loss_euclidean_ade = mean_n min_s mean_t || recon[s,n,t] - gt[n,t] ||_2
//   loss_euclidean_fde = mean_n min_s        || recon[s,n,T-1] - gt[n,T-1] ||_2
// One thread per pedestrian scans the S samples (torch.min semantics: first minimum, NaN wins), records the three
// arg-mins for the backward pass and its three minima; the last block to finish sums the per-pedestrian minima in
// index order (deterministic) and writes the means.  workspace: one uint32 ticket (zero on entry / exit).
__device__ __forceinline__ bool takes_min(float v, float best, bool first) {
  return first || v < best || (v != v && best == best);
}

__global__ void forward_losses_kernel(const float* __restrict__ C, const float* __restrict__ anchor_m,
                                      const float* __restrict__ anchor_s, const unsigned char* __restrict__ moving,
                                      const float* __restrict__ C_gt, const float* __restrict__ recon,
                                      const float* __restrict__ gt, int64_t n, int s, int k, int t,
                                      float* __restrict__ per_ped /* 3 x N minima */, int32_t* __restrict__ argmins /* 3 x N */,
                                      float* __restrict__ losses /* 3 */, unsigned* __restrict__ ticket) {
  __shared__ unsigned is_last;
  __shared__ double red[3][128];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float* an = moving[i] ? anchor_m : anchor_s;
    float cg[ET_MAX_K];
    for (int j = 0; j < k; ++j) cg[j] = __ldg(C_gt + (int64_t)j * n + i);
    float b_ec = 0.f, b_ade = 0.f, b_fde = 0.f;
    int a_ec = 0, a_ade = 0, a_fde = 0;
    for (int si = 0; si < s; ++si) {
      float ss = 0.f;
      for (int j = 0; j < k; ++j) {
        float v = __ldg(C + ((int64_t)j * n + i) * s + si);
        if (an) v = __ldg(an + j * s + si) + v;
        const float df = v - cg[j];
        ss = fmaf(df, df, ss);
      }
      const float ec = sqrtf(ss);
      const float2* rp = reinterpret_cast<const float2*>(recon) + ((int64_t)si * n + i) * t;
      const float2* gp = reinterpret_cast<const float2*>(gt) + i * t;
      float sum = 0.f, last = 0.f;
      for (int q = 0; q < t; ++q) {
        const float2 a = __ldg(rp + q), b = __ldg(gp + q);
        const float dx = a.x - b.x, dy = a.y - b.y;
        last = sqrtf(fmaf(dy, dy, dx * dx));
        sum += last;
      }
      const float ade = sum / (float)t;
      if (takes_min(ec, b_ec, si == 0)) { b_ec = ec; a_ec = si; }
      if (takes_min(ade, b_ade, si == 0)) { b_ade = ade; a_ade = si; }
      if (takes_min(last, b_fde, si == 0)) { b_fde = last; a_fde = si; }
    }
    per_ped[i] = b_ec; per_ped[n + i] = b_ade; per_ped[2 * n + i] = b_fde;
    argmins[i] = a_ec; argmins[n + i] = a_ade; argmins[2 * n + i] = a_fde;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int m = 0; m < 3; ++m) {
    double acc = 0.0;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) acc += (double)__ldcg(per_ped + m * n + e);
    red[m][threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double tot = 0.0;
    for (int q = 0; q < (int)blockDim.x; ++q) tot += red[threadIdx.x][q];
    losses[threadIdx.x] = (float)(tot / (double)n);
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// Gradient of w_ec * loss_eigentraj + w_ade * loss_ade + w_fde * loss_fde wrt C_refine (k,N,S): only the arg-min samples of
// each pedestrian receive a gradient.  The displacement part goes through d recon / d C = U_g^T ((. R) / sca).
// accumulate = 0: the thread first zero-fills its pedestrian's (k x S) block; 1: grad_C already holds a dense part.
__global__ void forward_losses_bwd_kernel(const float* __restrict__ C, const float* __restrict__ anchor_m,
                                          const float* __restrict__ anchor_s, const unsigned char* __restrict__ moving,
                                          const float* __restrict__ C_gt, const float* __restrict__ recon,
                                          const float* __restrict__ gt, int64_t n, int s, int k, int t2,
                                          const float* __restrict__ U_m, const float* __restrict__ U_s,
                                          const float* __restrict__ rot, const float* __restrict__ sca,
                                          const int32_t* __restrict__ argmins, const float* __restrict__ w /* 3 */,
                                          int accumulate, float* __restrict__ grad_C) {
  extern __shared__ float Us[];
  float* Um = Us;
  float* Usx = Um + t2 * k;
  stage(Um, U_m, t2 * k);
  stage(Usx, U_s, t2 * k);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = t2 / 2;
  const bool mv = moving[i] != 0;
  const float* an = mv ? anchor_m : anchor_s;
  const float* Ub = mv ? Um : Usx;
  if (!accumulate)
    for (int j = 0; j < k; ++j)
      for (int si = 0; si < s; ++si) grad_C[((int64_t)j * n + i) * s + si] = 0.f;
  const float inv_n = 1.0f / (float)n;
  const float w_ec = __ldg(w) * inv_n, w_ade = __ldg(w + 1) * inv_n / (float)t, w_fde = __ldg(w + 2) * inv_n;
  // coefficient loss: d ||v|| / d v = v / ||v|| (0 at v = 0)
  {
    const int si = argmins[i];
    float v[ET_MAX_K], ss = 0.f;
    for (int j = 0; j < k; ++j) {
      float c = __ldg(C + ((int64_t)j * n + i) * s + si);
      if (an) c = __ldg(an + j * s + si) + c;
      v[j] = c - __ldg(C_gt + (int64_t)j * n + i);
      ss = fmaf(v[j], v[j], ss);
    }
    const float nrm = sqrtf(ss);
    const float sc = nrm > 0.f ? w_ec / nrm : 0.f;
    for (int j = 0; j < k; ++j) grad_C[((int64_t)j * n + i) * s + si] += sc * v[j];
  }
  // displacement losses: ADE touches every frame of its arg-min sample, FDE the last frame of its own
  const NormState st = load_norm_state(nullptr, rot, sca, i, ET_NORM_ROT | ET_NORM_SCA);
  const float inv_sca = 1.0f / st.sca;
  const float2* gp = reinterpret_cast<const float2*>(gt) + i * t;
  for (int part = 0; part < 2; ++part) {
    const int si = argmins[(part + 1) * n + i];
    const float wt = part ? w_fde : w_ade;
    const float2* rp = reinterpret_cast<const float2*>(recon) + ((int64_t)si * n + i) * t;
    float g[FMAX2T];
    for (int q = 0; q < t; ++q) {
      float gx = 0.f, gy = 0.f;
      if (part == 0 || q == t - 1) {
        const float2 a = __ldg(rp + q), b = __ldg(gp + q);
        const float dx = a.x - b.x, dy = a.y - b.y;
        const float d = sqrtf(fmaf(dy, dy, dx * dx));
        const float sc = d > 0.f ? wt / d : 0.f;
        gx = sc * dx;
        gy = sc * dy;
      }
      g[2 * q] = (gx * st.r00 + gy * st.r10) * inv_sca;
      g[2 * q + 1] = (gx * st.r01 + gy * st.r11) * inv_sca;
    }
    for (int j = 0; j < k; ++j) {
      float acc = 0.f;
      for (int r = 0; r < t2; ++r) acc = fmaf(Ub[r * k + j], g[r], acc);
      grad_C[((int64_t)j * n + i) * s + si] += acc;
    }
  }
}

static int fwd_check(const char* who, int64_t n, int t, int k) {
  if (n < 0) return fail(ET_ERR_BADARG, "%s: n < 0", who);
  if (t < 1 || t > ET_MAX_T) return fail(ET_ERR_UNSUPPORTED, "%s: T = %d outside [1, %d]", who, t, ET_MAX_T);
  if (k < 1 || k > ET_MAX_K) return fail(ET_ERR_UNSUPPORTED, "%s: k = %d outside [1, %d]", who, k, ET_MAX_K);
  return ET_OK;
}

}  // namespace et

using namespace et;

extern "C" {

int et_forward_project(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, const float* U_obs_m,
                       const float* U_obs_s, const float* U_pred_m, const float* U_pred_s, int k, float static_dist,
                       float* C_obs, float* C_pred, float* ori, float* rot, float* sca, unsigned char* moving,
                       et_stream_t stream) {
  int rc = fwd_check("et_forward_project", n, t_obs, k);
  if (rc) return rc;
  if (pred && (rc = fwd_check("et_forward_project", n, t_pred, k))) return rc;
  ET_REQUIRE(t_obs >= 3, ET_ERR_UNSUPPORTED, "et_forward_project: T_obs = %d < 3", t_obs);
  if (n == 0) return ET_OK;
  ET_REQUIRE(obs && U_obs_m && U_obs_s && C_obs && ori && rot && sca && moving, ET_ERR_BADARG, "et_forward_project: null pointer");
  ET_REQUIRE(!pred || (U_pred_m && U_pred_s && C_pred), ET_ERR_BADARG, "et_forward_project: pred given but U_pred_* / C_pred null");
  ET_REQUIRE(aligned16(rot), ET_ERR_ALIGN, "et_forward_project: rot must be 16-byte aligned");
  const size_t smem = (size_t)(2 * 2 * t_obs + (pred ? 2 * 2 * t_pred : 0)) * k * sizeof(float);
  forward_project_kernel<<<(unsigned)((n + 127) / 128), 128, smem, as_stream(stream)>>>(
      obs, pred, n, 2 * t_obs, 2 * t_pred, U_obs_m, U_obs_s, U_pred_m, U_pred_s, k, static_dist, C_obs, C_pred, ori, rot, sca,
      moving);
  return check_launch("forward_project_kernel");
}

int et_forward_reconstruct(const float* C, const float* anchor_m, const float* anchor_s, int64_t n, int s, int k, int t,
                           const float* U_m, const float* U_s, const unsigned char* moving, const float* ori,
                           const float* rot, const float* sca, float* out, et_stream_t stream) {
  int rc = fwd_check("et_forward_reconstruct", n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1, ET_ERR_BADARG, "et_forward_reconstruct: S = %d", s);
  if (n == 0) return ET_OK;
  ET_REQUIRE(C && U_m && U_s && moving && ori && rot && sca && out, ET_ERR_BADARG, "et_forward_reconstruct: null pointer");
  ET_REQUIRE((anchor_m == nullptr) == (anchor_s == nullptr), ET_ERR_BADARG, "et_forward_reconstruct: give both anchor sets or none");
  ET_REQUIRE(aligned16(rot), ET_ERR_ALIGN, "et_forward_reconstruct: rot must be 16-byte aligned");
  forward_reconstruct_kernel<<<(unsigned)((n * s + 127) / 128), 128, (size_t)2 * 2 * t * k * sizeof(float), as_stream(stream)>>>(
      C, anchor_m, anchor_s, n, s, k, 2 * t, U_m, U_s, moving, ori, rot, sca, out);
  return check_launch("forward_reconstruct_kernel");
}

int et_forward_reconstruct_bwd(const float* grad_out, int64_t n, int s, int k, int t, const float* U_m, const float* U_s,
                               const unsigned char* moving, const float* rot, const float* sca, float* grad_C,
                               et_stream_t stream) {
  int rc = fwd_check("et_forward_reconstruct_bwd", n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1, ET_ERR_BADARG, "et_forward_reconstruct_bwd: S = %d", s);
  if (n == 0) return ET_OK;
  ET_REQUIRE(grad_out && U_m && U_s && moving && rot && sca && grad_C, ET_ERR_BADARG, "et_forward_reconstruct_bwd: null pointer");
  ET_REQUIRE(aligned16(rot), ET_ERR_ALIGN, "et_forward_reconstruct_bwd: rot must be 16-byte aligned");
  forward_reconstruct_bwd_kernel<<<(unsigned)((n * s + 127) / 128), 128, (size_t)2 * 2 * t * k * sizeof(float),
                                   as_stream(stream)>>>(grad_out, n, s, k, 2 * t, U_m, U_s, moving, rot, sca, grad_C);
  return check_launch("forward_reconstruct_bwd_kernel");
}

int et_forward_losses(const float* C, const float* anchor_m, const float* anchor_s, const unsigned char* moving,
                      const float* C_gt, const float* recon, const float* gt, int64_t n, int s, int k, int t,
                      float* per_ped, int32_t* argmins, float* losses, void* workspace, et_stream_t stream) {
  int rc = fwd_check("et_forward_losses", n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1 && n >= 1, ET_ERR_BADARG, "et_forward_losses: needs S >= 1 and N >= 1 (the mean over N of an empty batch is undefined)");
  ET_REQUIRE(C && moving && C_gt && recon && gt && per_ped && argmins && losses && workspace, ET_ERR_BADARG,
             "et_forward_losses: null pointer");
  ET_REQUIRE((anchor_m == nullptr) == (anchor_s == nullptr), ET_ERR_BADARG, "et_forward_losses: give both anchor sets or none");
  forward_losses_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(
      C, anchor_m, anchor_s, moving, C_gt, recon, gt, n, s, k, t, per_ped, argmins, losses, reinterpret_cast<unsigned*>(workspace));
  return check_launch("forward_losses_kernel");
}

int et_forward_losses_bwd(const float* C, const float* anchor_m, const float* anchor_s, const unsigned char* moving,
                          const float* C_gt, const float* recon, const float* gt, int64_t n, int s, int k, int t,
                          const float* U_m, const float* U_s, const float* rot, const float* sca, const int32_t* argmins,
                          const float* loss_weights, int accumulate, float* grad_C, et_stream_t stream) {
  int rc = fwd_check("et_forward_losses_bwd", n, t, k);
  if (rc) return rc;
  ET_REQUIRE(s >= 1, ET_ERR_BADARG, "et_forward_losses_bwd: S = %d", s);
  if (n == 0) return ET_OK;
  ET_REQUIRE(C && moving && C_gt && recon && gt && U_m && U_s && rot && sca && argmins && loss_weights && grad_C, ET_ERR_BADARG,
             "et_forward_losses_bwd: null pointer");
  ET_REQUIRE(aligned16(rot), ET_ERR_ALIGN, "et_forward_losses_bwd: rot must be 16-byte aligned");
  forward_losses_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, (size_t)2 * 2 * t * k * sizeof(float), as_stream(stream)>>>(
      C, anchor_m, anchor_s, moving, C_gt, recon, gt, n, s, k, 2 * t, U_m, U_s, rot, sca, argmins, loss_weights, accumulate, grad_C);
  return check_launch("forward_losses_bwd_kernel");
}

}  // extern "C"
