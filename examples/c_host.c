/*
 * A plain-C host of libet_b200.so: the eigen-basis (ETDescriptor.parameter_initialization, descriptor.py:116-142) and the
 * rank-k round trip (descriptor.py:144-176) of a batch of synthetic pedestrians, with nothing but the CUDA runtime and
 * include/et_b200.h -- no Python, no torch.  Build (see the Makefile target `example`):
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_host.c -o build/c_host \
 *       -Leigentrajectory_b200 -let_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,'$ORIGIN/../eigentrajectory_b200'
 *
 * Exit code 0: round trip within 5 % of the data (a rank-6 basis keeps > 95 % of a smooth 20-frame walk);
 * 2: no CUDA device; 1: a call failed (the message of et_last_error() is printed).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "et_b200.h"

#define CHECK_ET(call)                                                              \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != ET_OK) {                                                             \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, et_last_error());               \
      return 1;                                                                     \
    }                                                                               \
  } while (0)
#define CHECK_CUDA(call)                                                            \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      fprintf(stderr, "%s -> %s\n", #call, cudaGetErrorString(e_));                 \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

static double uniform(unsigned long long* state) {          /* 53-bit LCG draw in [0, 1) */
  *state = *state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(*state >> 11) / 9007199254740992.0;
}

int main(void) {
  enum { T_OBS = 8, T_PRED = 12, K = 6 };
  const int64_t n = 8192;
  const int flags = ET_NORM_ORI | ET_NORM_ROT | ET_NORM_SCA;
  int devices = 0;
  printf("libet_b200 version %d (header %d)\n", et_version(), ET_B200_VERSION);
  if (cudaGetDeviceCount(&devices) != cudaSuccess || devices == 0) {
    fprintf(stderr, "no CUDA device: the library has no CPU path\n");
    return 2;
  }

  /* synthetic walkers: constant speed, slowly turning heading, a little noise */
  float* obs_h = (float*)malloc(sizeof(float) * n * T_OBS * 2);
  float* pred_h = (float*)malloc(sizeof(float) * n * T_PRED * 2);
  unsigned long long seed = 12345;
  for (int64_t i = 0; i < n; ++i) {
    double x = 20.0 * uniform(&seed) - 10.0, y = 20.0 * uniform(&seed) - 10.0;
    double th = 6.283185307179586 * uniform(&seed), v = 0.2 + 0.6 * uniform(&seed), w = 0.1 * (uniform(&seed) - 0.5);
    for (int t = 0; t < T_OBS + T_PRED; ++t) {
      x += v * cos(th + w * t) + 0.03 * (uniform(&seed) - 0.5);
      y += v * sin(th + w * t) + 0.03 * (uniform(&seed) - 0.5);
      float* dst = t < T_OBS ? obs_h + (i * T_OBS + t) * 2 : pred_h + (i * T_PRED + (t - T_OBS)) * 2;
      dst[0] = (float)x;
      dst[1] = (float)y;
    }
  }

  float *obs, *pred, *rec_obs, *rec_pred, *C_obs, *C_pred, *U_obs, *U_pred, *S_obs, *S_pred;
  double* G;
  void* ws;
  const size_t ws_bytes = et_gram_workspace_bytes();
  CHECK_CUDA(cudaMalloc((void**)&obs, sizeof(float) * n * T_OBS * 2));
  CHECK_CUDA(cudaMalloc((void**)&pred, sizeof(float) * n * T_PRED * 2));
  CHECK_CUDA(cudaMalloc((void**)&rec_obs, sizeof(float) * n * T_OBS * 2));
  CHECK_CUDA(cudaMalloc((void**)&rec_pred, sizeof(float) * n * T_PRED * 2));
  CHECK_CUDA(cudaMalloc((void**)&C_obs, sizeof(float) * K * n));
  CHECK_CUDA(cudaMalloc((void**)&C_pred, sizeof(float) * K * n));
  CHECK_CUDA(cudaMalloc((void**)&U_obs, sizeof(float) * 2 * T_OBS * K));
  CHECK_CUDA(cudaMalloc((void**)&U_pred, sizeof(float) * 2 * T_PRED * K));
  CHECK_CUDA(cudaMalloc((void**)&S_obs, sizeof(float) * K));
  CHECK_CUDA(cudaMalloc((void**)&S_pred, sizeof(float) * K));
  CHECK_CUDA(cudaMalloc((void**)&G, sizeof(double) * (16 * 16 + 24 * 24)));
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  CHECK_CUDA(cudaMemset(G, 0, sizeof(double) * (16 * 16 + 24 * 24)));   /* et_gram ACCUMULATES (row-sharded callers sum shards) */
  CHECK_CUDA(cudaMemset(ws, 0, ws_bytes));                               /* barrier counters: zero on entry, zero again on exit */
  CHECK_CUDA(cudaMemcpy(obs, obs_h, sizeof(float) * n * T_OBS * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(pred, pred_h, sizeof(float) * n * T_PRED * 2, cudaMemcpyHostToDevice));

  /* basis: one pass for both float64 Gram matrices, one launch for both eigen-solves */
  CHECK_ET(et_gram(obs, pred, n, T_OBS, T_PRED, flags, G, G + 16 * 16, ws, 0));
  CHECK_ET(et_eig_jacobi_pair(G, 2 * T_OBS, G + 16 * 16, 2 * T_PRED, K, U_obs, S_obs, U_pred, S_pred, 0));
  /* rank-K round trip of every pedestrian (normalise -> project -> reconstruct -> denormalise), coefficients kept */
  CHECK_ET(et_project_reconstruct(obs, pred, n, T_OBS, T_PRED, U_obs, U_pred, K, flags, rec_obs, rec_pred, C_obs, C_pred, 0, 0));
  CHECK_CUDA(cudaDeviceSynchronize());

  float S_h[K];
  float* rec_h = (float*)malloc(sizeof(float) * n * T_PRED * 2);
  CHECK_CUDA(cudaMemcpy(S_h, S_pred, sizeof(S_h), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(rec_h, rec_pred, sizeof(float) * n * T_PRED * 2, cudaMemcpyDeviceToHost));
  double err = 0.0, ref = 0.0;
  for (int64_t i = 0; i < n; ++i)
    for (int t = 0; t < T_PRED; ++t)
      for (int c = 0; c < 2; ++c) {
        const int64_t e = (i * T_PRED + t) * 2 + c;
        /* measured relative to the last observed position: what the normaliser removes */
        const double d = pred_h[e] - obs_h[(i * T_OBS + T_OBS - 1) * 2 + c];
        err += (double)(rec_h[e] - pred_h[e]) * (rec_h[e] - pred_h[e]);
        ref += d * d;
      }
  printf("singular values of the prediction basis:");
  for (int j = 0; j < K; ++j) printf(" %.4g", S_h[j]);
  printf("\nrank-%d round trip of the futures: relative error %.3e (%lld kernels launched)\n", K, sqrt(err / ref),
         (long long)et_launch_count());
  return sqrt(err / ref) < 0.05 ? 0 : 1;
}
