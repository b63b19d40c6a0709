// Shared host/device helpers for libet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/et_b200.h"

namespace et {

// ---- host side -----------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);     // cudaGetLastError -> ET_OK / ET_ERR_CUDA; counts the launch
int sm_count();                          // SMs of the current device (cached per device)
int tune_get(int knob);                  // current value of an ET_TUNE_* knob (0 = default)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(et_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define ET_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) return ::et::fail((code), __VA_ARGS__); \
  } while (0)

// ---- device side ---------------------------------------------------------------------
// Per-pedestrian normaliser state (normalizer.py:17-28).  The rotation is kept as a general
// 2x2 matrix because TrajNorm.set_params (normalizer.py:36-40) accepts arbitrary state.
struct NormState {
  float ox, oy;               // origin  = last observed frame
  float r00, r01, r10, r11;   // [[c,-s],[s,c]], heading of (last - third_last)
  float sca;                  // (1/||d||) * 2
};

// last = obs[T-1], third = obs[T-3].  c,s = d/||d|| equals cos/sin(atan2(dy,dx)) of the
// reference to <= 3e-7 absolute; the zero vector maps to angle 0 as atan2(0,0) does.
__device__ __forceinline__ NormState make_norm_state(float lx, float ly, float tx, float ty) {
  NormState p;
  p.ox = lx;
  p.oy = ly;
  const float dx = lx - tx, dy = ly - ty;
  const float n2 = dx * dx + dy * dy;
  const float nrm = sqrtf(n2);
  float c = 1.f, s = 0.f;
  if (n2 > 0.f) {
    const float inv = 1.0f / nrm;
    c = dx * inv;
    s = dy * inv;
  }
  p.r00 = c; p.r01 = -s; p.r10 = s; p.r11 = c;
  p.sca = (1.0f / nrm) * 2.0f;   // inf when the pedestrian did not move, as in the reference
  return p;
}

// (a,b) <- ((a,b) - ori) @ R * sca
__device__ __forceinline__ void norm_fwd(float& a, float& b, const NormState& p, int flags) {
  if (flags & ET_NORM_ORI) { a -= p.ox; b -= p.oy; }
  if (flags & ET_NORM_ROT) {
    const float na = a * p.r00 + b * p.r10;
    const float nb = a * p.r01 + b * p.r11;
    a = na; b = nb;
  }
  if (flags & ET_NORM_SCA) { a *= p.sca; b *= p.sca; }
}

// (a,b) <- ((a,b) / sca) @ R^T + ori ; inv_sca = 1/sca precomputed by the caller (one IEEE
// division per pedestrian instead of one per coordinate).
__device__ __forceinline__ void norm_bwd(float& a, float& b, const NormState& p, float inv_sca, int flags) {
  if (flags & ET_NORM_SCA) { a *= inv_sca; b *= inv_sca; }
  if (flags & ET_NORM_ROT) {
    const float na = a * p.r00 + b * p.r01;
    const float nb = a * p.r10 + b * p.r11;
    a = na; b = nb;
  }
  if (flags & ET_NORM_ORI) { a += p.ox; b += p.oy; }
}

// Denormalisation as one affine map per pedestrian: out = (a, b) M + o with M = R^T / sca and o = ori folded once, so the
// per-point cost is four FMAs and no flag tests (differs from scale-then-rotate-then-shift by <= 1 ulp per operation).
struct AffineBwd {
  float m00, m01, m10, m11, ox, oy;
};
__device__ __forceinline__ AffineBwd make_affine_bwd(const NormState& p, int flags) {
  const float inv = (flags & ET_NORM_SCA) ? 1.0f / p.sca : 1.0f;
  AffineBwd t;
  if (flags & ET_NORM_ROT) { t.m00 = p.r00 * inv; t.m01 = p.r01 * inv; t.m10 = p.r10 * inv; t.m11 = p.r11 * inv; }
  else { t.m00 = inv; t.m01 = 0.f; t.m10 = 0.f; t.m11 = inv; }
  t.ox = (flags & ET_NORM_ORI) ? p.ox : 0.f;
  t.oy = (flags & ET_NORM_ORI) ? p.oy : 0.f;
  return t;
}
__device__ __forceinline__ void affine_bwd(float& a, float& b, const AffineBwd& t) {
  const float na = fmaf(a, t.m00, fmaf(b, t.m01, t.ox));
  const float nb = fmaf(a, t.m10, fmaf(b, t.m11, t.oy));
  a = na; b = nb;
}

// Normalisation likewise: out = ((a, b) - o) F with F = R * sca (two subtractions + four multiply-adds per point).
struct AffineFwd {
  float f00, f01, f10, f11, ox, oy;
};
__device__ __forceinline__ AffineFwd make_affine_fwd(const NormState& p, int flags) {
  const float sc = (flags & ET_NORM_SCA) ? p.sca : 1.0f;
  AffineFwd t;
  if (flags & ET_NORM_ROT) { t.f00 = p.r00 * sc; t.f01 = p.r01 * sc; t.f10 = p.r10 * sc; t.f11 = p.r11 * sc; }
  else { t.f00 = sc; t.f01 = 0.f; t.f10 = 0.f; t.f11 = sc; }
  t.ox = (flags & ET_NORM_ORI) ? p.ox : 0.f;
  t.oy = (flags & ET_NORM_ORI) ? p.oy : 0.f;
  return t;
}
__device__ __forceinline__ void affine_fwd(float& a, float& b, const AffineFwd& t) {
  const float da = a - t.ox, db = b - t.oy;
  a = fmaf(da, t.f00, db * t.f10);
  b = fmaf(da, t.f01, db * t.f11);
}

// Read a stored state (ori (N,1,2), rot (N,2,2), sca (N,1,1)); absent parts are identity.
__device__ __forceinline__ NormState load_norm_state(const float* ori, const float* rot, const float* sca,
                                                     int64_t i, int flags) {
  NormState p;
  p.ox = 0.f; p.oy = 0.f; p.r00 = 1.f; p.r01 = 0.f; p.r10 = 0.f; p.r11 = 1.f; p.sca = 1.f;
  if ((flags & ET_NORM_ORI) && ori) {
    const float2 o = __ldg(reinterpret_cast<const float2*>(ori) + i);
    p.ox = o.x; p.oy = o.y;
  }
  if ((flags & ET_NORM_ROT) && rot) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(rot) + i);
    p.r00 = r.x; p.r01 = r.y; p.r10 = r.z; p.r11 = r.w;
  }
  if ((flags & ET_NORM_SCA) && sca) p.sca = __ldg(sca + i);
  return p;
}

__device__ __forceinline__ void store_norm_state(float* ori, float* rot, float* sca, int64_t i,
                                                 const NormState& p, int flags) {
  if ((flags & ET_NORM_ORI) && ori) reinterpret_cast<float2*>(ori)[i] = make_float2(p.ox, p.oy);
  if ((flags & ET_NORM_ROT) && rot) reinterpret_cast<float4*>(rot)[i] = make_float4(p.r00, p.r01, p.r10, p.r11);
  if ((flags & ET_NORM_SCA) && sca) sca[i] = p.sca;
}

// Streaming 128-bit global accesses: data is touched once, keep it out of L1.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: two independent IEEE fp32 operations per instruction) -------
typedef unsigned long long f32x2_t;   // lo = bits 0..31 (lower address when loaded from memory), hi = bits 32..63
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ---- grid-wide reduction of per-block partial records (cooperative launches only) --------------
// One-shot, self-resetting barrier over all blocks of the grid.  ctr[0] = arrivals, ctr[1] = departures; both are
// zero on entry and zero again on exit.  Needs every block co-resident: launch with launch_cooperative().
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&ctr[0], 1u);
    while (*reinterpret_cast<volatile unsigned*>(&ctr[0]) < nblocks) __nanosleep(20);
    __threadfence();
    if (atomicAdd(&ctr[1], 1u) == nblocks - 1) {   // everybody has left the spin loop
      ctr[0] = 0u;
      ctr[1] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
}

// Sum element e over n_part partial records (stride rec doubles) with one warp: lanes stride over the
// partials, fixed shuffle tree => the result depends only on (n_part, data), never on timing.
__device__ __forceinline__ double warp_fold(const double* partials, int rec, int n_part, int e, int lane) {
  // four independent load/accumulate streams per lane: the L2 round trips overlap instead of forming one chain
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int p = lane;
  for (; p + 96 < n_part; p += 128) {
    const double v0 = __ldcg(partials + (size_t)p * rec + e);
    const double v1 = __ldcg(partials + (size_t)(p + 32) * rec + e);
    const double v2 = __ldcg(partials + (size_t)(p + 64) * rec + e);
    const double v3 = __ldcg(partials + (size_t)(p + 96) * rec + e);
    s0 += v0; s1 += v1; s2 += v2; s3 += v3;
  }
  for (; p < n_part; p += 32) s0 += __ldcg(partials + (size_t)p * rec + e);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Same for partials stored entry-major (v[0 .. n_part) contiguous for one entry): a warp's loads are coalesced and all
// in flight together, so the fold costs about one L2 round trip instead of one per 128 partial records.
__device__ __forceinline__ double warp_fold_contig(const double* v, int n_part, int lane) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int p = lane;
  for (; p + 96 < n_part; p += 128) {
    const double v0 = __ldcg(v + p);
    const double v1 = __ldcg(v + p + 32);
    const double v2 = __ldcg(v + p + 64);
    const double v3 = __ldcg(v + p + 96);
    s0 += v0; s1 += v1; s2 += v2; s3 += v3;
  }
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  if (p < n_part) t0 = __ldcg(v + p);
  if (p + 32 < n_part) t1 = __ldcg(v + p + 32);
  if (p + 64 < n_part) t2 = __ldcg(v + p + 64);
  s0 += t0; s1 += t1; s2 += t2;
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// NE entries at once (entry e of the group starts at v + e * stride): all NE * ceil(n_part / 32) loads of a lane are in
// flight together and the NE shuffle trees interleave, so a warp that has to fold many entries pays about one L2 round
// trip per group instead of one per entry.  Per entry the additions are exactly those of warp_fold_contig (same bits).
template <int NE>
__device__ __forceinline__ void warp_fold_contig_multi(const double* v, size_t stride, int n_valid, int n_part, int lane,
                                                       double (&out)[NE]) {
  double s0[NE], s1[NE], s2[NE], s3[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) s0[e] = s1[e] = s2[e] = s3[e] = 0.0;
  int p = lane;
  for (; p + 96 < n_part; p += 128) {
    double v0[NE], v1[NE], v2[NE], v3[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const double* ve = v + (size_t)(e < n_valid ? e : 0) * stride;
      v0[e] = __ldcg(ve + p);
      v1[e] = __ldcg(ve + p + 32);
      v2[e] = __ldcg(ve + p + 64);
      v3[e] = __ldcg(ve + p + 96);
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) { s0[e] += v0[e]; s1[e] += v1[e]; s2[e] += v2[e]; s3[e] += v3[e]; }
  }
  double t0[NE], t1[NE], t2[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const double* ve = v + (size_t)(e < n_valid ? e : 0) * stride;
    t0[e] = p < n_part ? __ldcg(ve + p) : 0.0;
    t1[e] = p + 32 < n_part ? __ldcg(ve + p + 32) : 0.0;
    t2[e] = p + 64 < n_part ? __ldcg(ve + p + 64) : 0.0;
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    s0[e] += t0[e]; s1[e] += t1[e]; s2[e] += t2[e];
    out[e] = (s0[e] + s1[e]) + (s2[e] + s3[e]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int e = 0; e < NE; ++e) out[e] += __shfl_xor_sync(0xffffffffu, out[e], o);
  }
}

// Cooperative launch wrapper (guarantees co-residency or fails loudly).
template <typename... Args>
inline cudaError_t launch_cooperative(void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args... args) {
  void* argv[] = {(void*)&args...};
  return cudaLaunchCooperativeKernel((const void*)kernel, grid, block, argv, smem, stream);
}

}  // namespace et
