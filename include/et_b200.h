/*
 * et_b200.h -- C ABI of libet_b200.so, the B200 (sm_100a) implementation of the
 * EigenTrajectory descriptor hot path.
 *
 * The reference (InhwanBae/EigenTrajectory) is pure Python/PyTorch and has no FFI of
 * its own; the entry points below are what a binding for this path attaches to.  Each
 * one cites the reference code it replaces (path:line relative to the reference
 * checkout).  The Python classes in eigentrajectory_b200/ (ETDescriptor, ETAnchor,
 * TrajNorm, BatchKMeans, compute_batch_ade/fde) call these through ctypes and nothing
 * else; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the library never
 *    allocates, frees or synchronises: the caller owns all buffers and the stream;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *  - return value: 0 (ET_OK) or a negative ET_ERR_* code; et_last_error() returns a
 *    thread-local, human readable message for the last failure on this thread;
 *  - trajectories are float32, contiguous (N, T, 2) "NTC"; coefficient matrices are
 *    float32 contiguous (k, N) / (k, N, S) exactly as the reference lays them out;
 *  - float4-vectorised and TMA paths need 16-byte aligned base pointers (ET_ERR_ALIGN
 *    otherwise); torch allocations always are.
 *  - normaliser flags are a bit-or of ET_NORM_ORI | ET_NORM_ROT | ET_NORM_SCA
 *    (TrajNorm(ori, rot, sca), EigenTrajectory/normalizer.py:13-15).
 */
#ifndef ET_B200_H_
#define ET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ET_B200_VERSION 100

enum {
  ET_OK = 0,
  ET_ERR_BADARG = -1,      /* null pointer, negative size, unsupported shape       */
  ET_ERR_ALIGN = -2,       /* pointer not 16-byte aligned                           */
  ET_ERR_CUDA = -3,        /* a CUDA runtime / driver call or a launch failed       */
  ET_ERR_UNSUPPORTED = -4, /* shape outside what the kernels were built for         */
  ET_ERR_NCCL = -5         /* libnccl could not be loaded or a collective call failed */
};

enum { ET_NORM_ORI = 1, ET_NORM_ROT = 2, ET_NORM_SCA = 4 };

/* Limits of the generic kernels (fast paths exist for T_obs=8, T_pred=12, k=6, S=20). */
#define ET_MAX_T 32    /* frames per trajectory segment (2T <= 64 coordinates)            */
#define ET_MAX_K 32    /* rank of the eigen-basis                                         */
#define ET_MAX_CLUSTERS 64
#define ET_MAX_KM_DIM 16

typedef void* et_stream_t;

/* ---- library ---------------------------------------------------------------------- */
int et_version(void);
const char* et_last_error(void);
/* Device the calling thread is on: SM count and compute capability. */
int et_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Total number of kernels this library has launched from this process (all threads). */
int64_t et_launch_count(void);

/* Plumbing for host-buffer callers: asynchronous pitched copy (either direction, pinned host memory
 * for true asynchrony) of `rows` rows of width_bytes, e.g. a (k, chunk) slice of a (k, N) matrix. */
int et_memcpy_2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                       size_t width_bytes, size_t rows, et_stream_t stream);

/* Launch-shape knobs for performance experiments (process-wide; 0 = the shipped default).
 * Results never depend on them beyond the last bits.
 * ET_TUNE_EIG_THREADS: 32 / 96 / 128 / 288 = the four-barrier Jacobi body with that many threads; 2001 / 2002 / 2003 =
 * first generation / second generation / second generation with the symmetric (upper-triangle) update of the two-barrier
 * body for the 16 x 16 and 24 x 24 bases (default: 2003; the environment variable ET_EIG_GEN = 1 | 2 | 3 sets the default
 * of the process); 3001 / 3002 / 3003 = the same with phase cycle counters written to info[2..13] of et_eig_jacobi
 * (diagnostic: info must then hold 14 ints). */
enum { ET_TUNE_ADE_CONFIG = 0, ET_TUNE_REC_BLOCKS_PER_SM = 1, ET_TUNE_GRAM_UNROLL = 2, ET_TUNE_EIG_THREADS = 3,
       ET_TUNE_PDL = 4, ET_TUNE_SEED_STEPWISE = 5 /* 1: farthest-point seeding as one launch per step */,
       ET_TUNE_KM_TOURNAMENT = 6 /* 1: per-group tournament arg-max instead of the ascending compare/select scan (A/B) */,
       ET_TUNE_KM_WARPS = 7 /* 16: sixteen warps per SM on shared record columns instead of twelve on private ones (A/B) */,
       ET_TUNE_COUNT = 8 };
int et_tune(int knob, int value);

/* ---- normaliser: EigenTrajectory/normalizer.py ------------------------------------- */
/* TrajNorm.calculate_params (normalizer.py:17-28).  obs (N,T_obs,2), T_obs >= 3.
 * ori (N,1,2), rot (N,2,2) = [[c,-s],[s,c]], sca (N,1,1) = (1/||d||)*2; outputs whose
 * flag bit is clear may be null and are not written. */
int et_norm_params(const float* obs, int64_t n, int t_obs, int flags,
                   float* ori, float* rot, float* sca, et_stream_t stream);
/* TrajNorm.normalize (normalizer.py:42-51): out = ((x - ori) @ R) * sca. */
int et_normalize(const float* traj, int64_t n, int t, int flags, const float* ori,
                 const float* rot, const float* sca, float* out, et_stream_t stream);
/* TrajNorm.denormalize (normalizer.py:53-62): out = ((x / sca) @ R^T) + ori. */
int et_denormalize(const float* traj, int64_t n, int t, int flags, const float* ori,
                   const float* rot, const float* sca, float* out, et_stream_t stream);

/* ---- descriptor: EigenTrajectory/descriptor.py -------------------------------------- */
/* ETDescriptor.to_ET_space (descriptor.py:59-73; same body at anchor.py:22-36):
 * C (k,N) = U^T M with M the (2T,N) view of traj (N,T,2); U (2T,k) row-major. */
int et_to_et_space(const float* traj, int64_t n, int t, const float* U, int k, float* C,
                   et_stream_t stream);
/* ETDescriptor.to_Euclidean_space (descriptor.py:75-89): traj (N,T,2) = (U C)^T.
 * C is addressed as C[j*ldc_k + i*ldc_n], so a (k,N,S) slice [:, :, s] is ldc_k=N*S,
 * ldc_n=S with the base pointer advanced by s. */
int et_to_euclidean_space(const float* C, int64_t ldc_k, int64_t ldc_n, int64_t n, int t,
                          const float* U, int k, float* traj, et_stream_t stream);
/* ETDescriptor.projection (descriptor.py:144-160) fused with normalize_trajectory
 * (descriptor.py:29-45): derives the normaliser state from obs, writes it (ori/rot/sca,
 * see et_norm_params), and writes C_obs (k,N) and, when pred != null, C_pred (k,N). */
int et_project(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred,
               const float* U_obs, const float* U_pred, int k, int flags, float* C_obs,
               float* C_pred, float* ori, float* rot, float* sca, et_stream_t stream);
/* ETDescriptor.reconstruction (descriptor.py:162-176) fused with ETAnchor.forward
 * (anchor.py:76-88): out (S,N,T,2) = denormalise(U (C[:,:,s] + anchor[:,s])).
 * C (k,N,S); anchor (k,S) or null; state as written by et_project. */
int et_reconstruct(const float* C, const float* anchor, int64_t n, int s, int k, int t,
                   const float* U, int flags, const float* ori, const float* rot,
                   const float* sca, float* out, et_stream_t stream);
/* Gradient of et_reconstruct wrt C (autograd through descriptor.py:173-175; U, anchor
 * and the state are constants, descriptor.py:87, anchor.py:87):
 * grad_C (k,N,S) = U^T (R^T-rotate(grad_out[s,n]) / sca). */
int et_reconstruct_bwd(const float* grad_out, int64_t n, int s, int k, int t, const float* U,
                       int flags, const float* rot, const float* sca, float* grad_C,
                       et_stream_t stream);
/* Headline op (BASELINE.json config 2): rank-k round trip of obs and pred in one pass,
 * the S=1 shape of script/descriptor_evaluation.py:94-107 on top of descriptor.py:144-176.
 * rec_obs (N,T_obs,2), rec_pred (N,T_pred,2); C_obs / C_pred (k,N) are optional (null =
 * coefficients are not materialised: 320 instead of 368 algorithmic bytes / trajectory).
 * variant: 0 = auto, 1 = direct global access kernel, 2..4 = persistent warp-specialised
 * TMA-tiled kernel (2: 4-stage ring, 2 blocks/SM, coefficients written by tensor stores when
 * N % 4 == 0; 3: 3-stage, 3 blocks/SM; 4: as 2 but coefficients stored by the consumer warps);
 * the TMA variants need (T_obs,T_pred,k) = (8,12,6) and n < 2^31. */
int et_project_reconstruct(const float* obs, const float* pred, int64_t n, int t_obs,
                           int t_pred, const float* U_obs, const float* U_pred, int k,
                           int flags, float* rec_obs, float* rec_pred, float* C_obs,
                           float* C_pred, int variant, et_stream_t stream);

/* ---- EigenTrajectory.forward glue without boolean-mask gathers (model.py:73-105) -------------- */
/* Per pedestrian: moving iff ||(last - third_last)/2|| > static_dist (model.py:46,73); moving rows are
 * normalised with ori|rot|sca and projected on U_*_m, static rows with ori|rot on U_*_s -- what
 * ET_m_descriptor / ET_s_descriptor.projection do on the two gathered groups (model.py:80-83).
 * Outputs: C_obs (k,N), C_pred (k,N) when pred != null, the normaliser state of every row
 * (ori (N,1,2), rot (N,2,2), sca (N,1,1) with sca = 1 on static rows) and moving (N) uint8. */
int et_forward_project(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred,
                       const float* U_obs_m, const float* U_obs_s, const float* U_pred_m,
                       const float* U_pred_s, int k, float static_dist, float* C_obs, float* C_pred,
                       float* ori, float* rot, float* sca, unsigned char* moving, et_stream_t stream);
/* Anchor refinement + reconstruction of both groups in one launch (model.py:98-105):
 * out (S,N,T,2) = denormalise(U_g (C[:,n,:] + anchor_g)), g = moving[n].  Anchors (k,S): both or none. */
int et_forward_reconstruct(const float* C, const float* anchor_m, const float* anchor_s, int64_t n,
                           int s, int k, int t, const float* U_m, const float* U_s,
                           const unsigned char* moving, const float* ori, const float* rot,
                           const float* sca, float* out, et_stream_t stream);
/* Gradient of et_forward_reconstruct wrt C: grad_C (k,N,S) = U_g^T ((grad_out R) / sca). */
int et_forward_reconstruct_bwd(const float* grad_out, int64_t n, int s, int k, int t,
                               const float* U_m, const float* U_s, const unsigned char* moving,
                               const float* rot, const float* sca, float* grad_C, et_stream_t stream);

/* The three training losses of model.py:119-123 in one launch: losses[3] = {loss_eigentraj,
 * loss_euclidean_ade, loss_euclidean_fde} (means over N of per-pedestrian minima over S, torch.min
 * semantics); per_ped (3,N) minima and argmins (3,N) int32 are kept for the backward pass.
 * C (k,N,S) are the refined coefficients BEFORE the anchor add; C_gt (k,N); recon (S,N,T,2); gt (N,T,2).
 * workspace: 4 bytes, zero-filled once. */
int et_forward_losses(const float* C, const float* anchor_m, const float* anchor_s,
                      const unsigned char* moving, const float* C_gt, const float* recon,
                      const float* gt, int64_t n, int s, int k, int t, float* per_ped,
                      int32_t* argmins, float* losses, void* workspace, et_stream_t stream);
/* grad_C (k,N,S) of sum_m loss_weights[m] * losses[m] (loss_weights: DEVICE float[3], the upstream
 * gradients of the three scalars): only arg-min samples receive gradient; the displacement terms are
 * chained through d recon / d C.  accumulate = 0 overwrites grad_C, 1 adds to it. */
int et_forward_losses_bwd(const float* C, const float* anchor_m, const float* anchor_s,
                          const unsigned char* moving, const float* C_gt, const float* recon,
                          const float* gt, int64_t n, int s, int k, int t, const float* U_m,
                          const float* U_s, const float* rot, const float* sca,
                          const int32_t* argmins, const float* loss_weights, int accumulate,
                          float* grad_C, et_stream_t stream);

/* ---- eigen-basis: ETDescriptor.truncated_SVD (descriptor.py:91-114) ------------------ */
/* One pass over the data: G_obs (2T_obs x 2T_obs) += M_obs M_obs^T and, when pred != null,
 * G_pred += M_pred M_pred^T, in float64 (FP64 tensor-core DMMA on the (8,12) fast path), where
 * M_* are the normalised matrices (normalisation by `flags` fused; flags = 0 and pred = null
 * gives the plain Gram matrix of an already normalised (N,T,2) tensor).  Accumulates into G
 * (caller zeroes it); the sum over row shards of N is what a multi-GPU caller all-reduces.
 * workspace: et_gram_workspace_bytes() bytes, zero-filled once by the caller (the kernel leaves
 * its ticket word zero again, so the buffer can be reused without clearing). */
size_t et_gram_workspace_bytes(void);
int et_gram(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, int flags,
            double* G_obs, double* G_pred, void* workspace, et_stream_t stream);
/* The data pass of ETDescriptor.parameter_initialization (descriptor.py:116-135) in ONE launch on the (8, 12) fast
 * path: et_gram plus, from the same read of the trajectories, the normaliser state of every row (ori / rot / sca as
 * et_norm_params writes them; each optional) and the normalised futures pred_norm (N, T_pred, 2) (optional) that the
 * anchor step consumes.  Other shapes compose et_norm_params + et_normalize + et_gram. */
int et_gram_init(const float* obs, const float* pred, int64_t n, int t_obs, int t_pred, int flags,
                 double* G_obs, double* G_pred, float* pred_norm, float* ori, float* rot, float* sca,
                 void* workspace, et_stream_t stream);
/* Symmetric eigen-solve of G (m x m, float64, m <= 64) by parallel-ordered cyclic Jacobi;
 * returns the k leading left singular vectors U (m,k) row-major float32 and singular
 * values S (k) = sqrt(lambda).  Column signs are canonical: the largest-magnitude
 * component of each column is positive.  U64 (m,k) / S64 (k) optional float64 copies;
 * info (optional, device int32[2]) receives {sweeps executed, rotations applied}. */
int et_eig_jacobi(const double* G, int m, int k, float* U, float* S, double* U64, double* S64,
                  int* info, et_stream_t stream);
/* The two eigen-solves of one descriptor (ETDescriptor.parameter_initialization, descriptor.py:134-135: observation and
 * prediction bases) in one launch, side by side on two SMs; (m_a, m_b) = (16, 24) takes the fused kernel, any other
 * pair of sizes is two et_eig_jacobi launches.  Results are bit-identical to et_eig_jacobi. */
int et_eig_jacobi_pair(const double* G_a, int m_a, const double* G_b, int m_b, int k, float* U_a,
                       float* S_a, float* U_b, float* S_b, et_stream_t stream);
/* Batched small-N SVD: one-sided (Hestenes) Jacobi on shared-memory resident tall-skinny
 * matrices.  Problem b is the (n_b x 2T) matrix of the trajectories
 * traj[offsets[b] .. offsets[b+1]) (already normalised).  offsets is a DEVICE int64 array
 * of batch+1 entries; max_rows >= max_b n_b sizes shared memory (n_b*2T*4 bytes must fit
 * 200 KB).  U (batch, 2T, k), S (batch, k). */
int et_svd_small(const float* traj, const int64_t* offsets, int batch, int64_t max_rows, int t,
                 int k, float* U, float* S, et_stream_t stream);

/* ---- k-means: EigenTrajectory/kmeans.py ----------------------------------------------- */
/* BatchKMeans.get_labels (kmeans.py:143-158) fused with the accumulation half of
 * compute_centroids (kmeans.py:160-184).  data (l,d,N), centroids (l,d,K), d <= 16, K <= 64.
 * sim = fl(fl(fl(2*dot) - |a|^2) - |b|^2), dot an ascending fp32 FMA chain from 0 and the
 * squared norms summed in torch's CPU order (bit-equal to the reference's CPU result, see
 * DESIGN.md), arg-max with lowest index on ties, NaN wins.
 * Optional outputs: labels int64 (l,N), maxsims float (l,N).
 * Optional accumulators (float64, ADDED to, caller zeroes once): sums (l,d,K), counts (l,K),
 * simsum (l) = sum of maxsims.  workspace: et_kmeans_workspace_bytes() bytes, zero-filled once.
 * status (optional, device int32[2]): when status[0] != 0 the call is a no-op (see finalize). */
size_t et_kmeans_workspace_bytes(int l, int d, int k_clusters);
int et_kmeans_assign(const float* data, const float* centroids, int l, int d, int64_t n,
                     int k_clusters, int64_t* labels, float* maxsims, double* sums,
                     double* counts, double* simsum, void* workspace, const int32_t* status,
                     et_stream_t stream);
/* et_kmeans_assign on a ROW SHARD: data (l,d,n) holds the global columns [row_offset, row_offset + n) of a data set
 * with n_global columns.  The |a|^2 summation order of a column follows its GLOBAL index and the GLOBAL column count,
 * so labels and similarities are the same bits as et_kmeans_assign on the unsharded tensor would give for these
 * columns.  n_global < 2^32 - 2^21. */
int et_kmeans_assign_shard(const float* data, const float* centroids, int l, int d, int64_t n,
                           int k_clusters, int64_t* labels, float* maxsims, double* sums,
                           double* counts, double* simsum, void* workspace, const int32_t* status,
                           int64_t row_offset, int64_t n_global, et_stream_t stream);
/* The whole Lloyd loop of BatchKMeans.fit (kmeans.py:226-239) in ONE persistent cooperative launch: up to max_iter
 * iterations of {get_labels, compute_centroids, calculate_error, `if error <= tol: break`} with the grid-wide sums
 * folded behind in-kernel grid barriers, no host round trip and no relaunch (the data stay L2 resident), then one
 * more scan that writes the labels of the last assignment -- the ones fit() returns -- when labels != null.
 * centroids (l,d,K) in; centroids_out (l,d,K) = centroids after the last update; optional err[1] (float64, error of
 * the last iteration over all l), status (device int32[2]) = {converged, iterations done}, simsum_last (l) = sum of
 * the best similarities of the last assignment (inertia = -simsum_last / N).  Arithmetic identical to the
 * et_kmeans_assign + et_kmeans_finalize sequence.  workspace: et_kmeans_workspace_bytes(), zero-filled once. */
int et_kmeans_lloyd(const float* data, const float* centroids, int l, int d, int64_t n, int k_clusters,
                    int max_iter, double tol, float* centroids_out, int64_t* labels, double* err,
                    int32_t* status, double* simsum_last, void* workspace, et_stream_t stream);
/* et_kmeans_lloyd over ROW SHARDS on several GPUs of one node, with the per-iteration exchange fused into the kernel:
 * every rank runs the persistent kernel on its n_local rows; after the local grid fold it stores its folded record
 * (d*K + K + 1 doubles per batch entry) directly into every rank's exchange buffer through peer mappings
 * (NVLink / NVSwitch), raises a per-iteration flag there, waits for the other ranks' flags in its own buffer and adds
 * the `world` records in rank order -- the all-reduce of SURVEY section 8e inside the kernel: no NCCL call, no host
 * round trip, identical centroids / convergence decision on every rank by construction.
 * exchange_peers: DEVICE array of `world` pointers, entry r = rank r's exchange buffer as mapped in THIS process
 * (et_kmeans_exchange_bytes() bytes each, zero-filled once before the first call, e.g. a torch symmetric-memory
 * allocation).  stamp_base: a number that grows by at least max_iter + 2 from one call to the next on the same
 * buffers (the same on every rank).  labels / simsum_last cover the local rows / the global sum.  world <= 16, l <= 32;
 * n_local may be 0.  row_offset / n_global: global index of this shard's first column and the global column count
 * (as et_kmeans_assign_shard; n_global = 0: the shard is numbered on its own).  A rank that never arrives traps the
 * kernel after ~10 s instead of hanging it. */
size_t et_kmeans_exchange_bytes(int l, int d, int k_clusters, int world);
int et_kmeans_lloyd_sharded(const float* data, const float* centroids, int l, int d, int64_t n_local,
                            int k_clusters, int max_iter, double tol, float* centroids_out,
                            int64_t* labels, double* err, int32_t* status, double* simsum_last,
                            void* workspace, int rank, int world, void* const* exchange_peers,
                            unsigned stamp_base, int64_t row_offset, int64_t n_global, et_stream_t stream);
/* Accumulation half of compute_centroids (kmeans.py:160-184) for caller-supplied labels
 * (l,N) int64; labels outside [0,K) are ignored.  sums / counts / workspace as above. */
int et_kmeans_accumulate(const float* data, const int64_t* labels, int l, int d, int64_t n,
                         int k_clusters, double* sums, double* counts, void* workspace,
                         et_stream_t stream);
/* Division half of compute_centroids + calculate_error (kmeans.py:45-51,183):
 * new_centroids = float(sums / counts) (0/0 -> NaN as in the reference);
 * err[0] = sum((old - new)^2) in float64 when err and old_centroids are given; sums / counts
 * are cleared for the next iteration, and so is simsum (l) after being copied to simsum_last (l)
 * (both optional).  With status (device int32[2]): a no-op when status[0] != 0, otherwise
 * status[1] += 1 and status[0] = (err <= tol) -- the reference's `if error <= self.tol: break`
 * (kmeans.py:239) evaluated on the device, so that a host can enqueue several Lloyd iterations
 * without synchronising. */
int et_kmeans_finalize(double* sums, double* counts, int l, int d, int k_clusters,
                       const float* old_centroids, float* new_centroids, double* err, double tol,
                       int32_t* status, double* simsum, double* simsum_last, et_stream_t stream);
/* BatchKMeans.kmeanspp (kmeans.py:78-112): deterministic farthest-point seeding.
 * centroids (l,d,K) out; column 0 = data[..., first_index]; every later column is the point
 * whose best similarity to the columns chosen so far is lowest (lowest index on ties),
 * with the reference's arithmetic (the reference recomputes every similarity every step; here the running best of a
 * point is reused across the steps in which it cannot have changed a bit and recomputed in the others, see
 * et_kmeans.cu).  All K - 1 steps run in ONE persistent cooperative launch while the points fit the SMs' shared memory
 * (~1.3e6 six-dimensional points on a B200), otherwise one launch per step.  scratch: l*K + 16 uint64. */
int et_kmeans_farthest_init(const float* data, int l, int d, int64_t n, int k_clusters,
                            int64_t first_index, float* centroids, unsigned long long* scratch,
                            et_stream_t stream);
/* k-means++ seeding by D^2 sampling with greedy local trials -- what sklearn.cluster.KMeans(init="k-means++") does
 * inside ETAnchor.anchor_generation (anchor.py:65-71) -- in ONE persistent cooperative launch.  uniform (l, K, trials)
 * float64 in [0, 1) from the caller's random generator: [., 0, 0] selects the first centre (index floor(u * N)),
 * [., i, j] the j-th candidate of step i by inverting the cumulative sum of the squared distances to the nearest centre
 * chosen so far; of the `trials` candidates (1..8; sklearn: 2 + floor(ln K)) the one with the lowest resulting
 * potential is kept (lowest trial index on ties).  Deterministic for given random numbers (float64 sums in a fixed
 * order).  The points must fit the SMs' shared memory (~1.3e6 six-dimensional points per batch entry at l = 1),
 * otherwise ET_ERR_UNSUPPORTED.  centroids (l,d,K) out; workspace: et_kmeans_d2_workspace_bytes(l, trials). */
size_t et_kmeans_d2_workspace_bytes(int l, int trials);
int et_kmeans_d2_init(const float* data, int l, int d, int64_t n, int k_clusters, int trials,
                      const double* uniform, float* centroids, void* workspace, et_stream_t stream);
/* et_kmeans_farthest_init over ROW SHARDS on several GPUs of one node: data (l,d,n_local) holds the global columns
 * [row_offset, row_offset + n_local) of n_global; first_global_index picks column 0.  Per step every rank's candidate
 * (similarity bits, GLOBAL index, coordinates) is stored into every rank's exchange buffer through peer memory and the
 * smallest key wins on every rank -- the (value, index) arg-min exchange and the winner's broadcast inside the kernel.
 * exchange_peers / stamp_base as et_kmeans_lloyd_sharded (stamp_base must grow by at least K + 2 per call); centroids
 * (l,d,K) are identical on every rank and equal to the unsharded call's.  scratch: l*K + 16 uint64.  Returns
 * ET_ERR_UNSUPPORTED when the local points do not fit the resident kernel (the caller then seeds with
 * et_kmeans_seed_candidate / et_kmeans_seed_fetch + two small all-reduces per step). */
int et_kmeans_farthest_init_sharded(const float* data, int l, int d, int64_t n_local, int k_clusters,
                                    int64_t first_global_index, int64_t row_offset, int64_t n_global,
                                    float* centroids, unsigned long long* scratch, int rank, int world,
                                    void* const* exchange_peers, unsigned stamp_base, et_stream_t stream);

/* One farthest-point step on a ROW SHARD (kmeans.py:95-98 evaluated on local rows): with the first
 * ncols columns of centroids (l,d,K) given, key_out[l] = min over local points of
 * (order-preserving bits of the point's best similarity) << 32 | local index.  A multi-GPU caller
 * takes the minimum over ranks of (value, global index) and broadcasts the winner's coordinates. */
int et_kmeans_seed_step(const float* data, const float* centroids, int l, int d, int64_t n,
                        int k_clusters, int ncols, unsigned long long* key_out, et_stream_t stream);

/* The two halves of a row-sharded farthest-point step without host round trips: et_kmeans_seed_candidate =
 * et_kmeans_seed_step with the key made GLOBAL and signed-orderable (index + row_offset in the low 32 bits, top bit
 * flipped; an empty shard proposes INT64_MAX), so that ONE int64 MIN all-reduce over the ranks elects the winner;
 * et_kmeans_seed_fetch then writes the winner's coordinates (l,d) as float64 on the rank that owns the column and zeros
 * elsewhere, so that ONE SUM all-reduce hands them to every rank.  row_offset + n <= 2^32.  n_global (0: the shard is
 * numbered on its own): the |a|^2 summation order then follows the global column numbering (as et_kmeans_assign_shard). */
int et_kmeans_seed_candidate(const float* data, const float* centroids, int l, int d, int64_t n,
                             int k_clusters, int ncols, int64_t row_offset, int64_t n_global,
                             long long* gkey_out, et_stream_t stream);
int et_kmeans_seed_fetch(const float* data, int l, int d, int64_t n, int64_t row_offset,
                         const long long* gkey, double* coords, et_stream_t stream);

/* ---- collectives for row-sharded callers ------------------------------------------------
 * The reference has no multi-GPU path of its own; these are what a host other than the Python mirror uses between the
 * local passes: et_gram on the local rows -> et_allreduce_f64 on the packed Gram accumulators -> et_eig_jacobi_pair
 * (identical bases on every rank, no broadcast); per Lloyd iteration et_kmeans_assign_shard -> et_allreduce_f64 on
 * [sums | counts | simsum] -> et_kmeans_finalize; per seeding step et_kmeans_seed_candidate -> et_allreduce_min_i64 ->
 * et_kmeans_seed_fetch -> et_allreduce_f64.  One communicator per process and GPU (NCCL underneath, bound at run time
 * with dlopen("libnccl.so.2"); ET_ERR_NCCL if it is missing).  All calls are enqueued on `stream`, in place. */
#define ET_COMM_ID_BYTES 128
typedef struct et_comm* et_comm_t;
/* rank 0 creates the 128-byte id and hands it to the other ranks by any out-of-band means */
int et_comm_unique_id(void* id_out);
/* collective over all ranks; binds the CURRENT CUDA device of the calling thread to `rank` */
int et_comm_init(int rank, int nranks, const void* unique_id, et_comm_t* comm_out);
int et_comm_rank(et_comm_t comm, int* rank, int* nranks);
int et_allreduce_f64(double* buf, size_t n, et_comm_t comm, et_stream_t stream);          /* sum */
int et_allreduce_min_i64(long long* buf, size_t n, et_comm_t comm, et_stream_t stream);   /* signed minimum */
int et_comm_destroy(et_comm_t comm);

/* ---- metrics: utils/metrics.py ---------------------------------------------------------- */
/* compute_batch_ade + compute_batch_fde (metrics.py:73-102) in one pass, optionally with
 * compute_batch_tcc (metrics.py:105-130) from the same read of pred.
 * pred (S,N,T,2), gt (N,T,2); ade (N), fde (N) float; argmin_fde (N) int32 optional;
 * tcc (N) float optional (T >= 2). */
int et_ade_fde(const float* pred, const float* gt, int s, int64_t n, int t, float* ade,
               float* fde, int32_t* argmin_fde, float* tcc, et_stream_t stream);
/* compute_batch_col (metrics.py:133-155): all N pedestrians are one scene; col (N) = percent of
 * samples in which the pedestrian comes within `thres` (reference: 0.2) of another one during
 * the first 14 quarter-frame interpolated steps. */
int et_col(const float* pred, int s, int64_t n, int t, float thres, float* col, et_stream_t stream);

/* ---- dataset preprocessing (HOST side): utils/dataloader.py ------------------------------ */
/* read_file (dataloader.py:121-132): parse a text buffer of `<frame><delim><ped><delim><x><delim><y>` lines
 * (each line stripped, split on delim, every field through float()) into rows_host (capacity, 4) float64.
 * rows_host = null only counts.  *n_rows_host receives the number of rows.  A blank line, a field that is not a
 * number or a line without exactly four fields is ET_ERR_BADARG (the reference raises ValueError / fails later). */
int et_dataset_parse_host(const char* text_host, size_t len, char delim, double* rows_host,
                          int64_t capacity, int64_t* n_rows_host);
/* The per-file body of TrajectoryDataset.__init__ (dataloader.py:189-226) plus poly_fit (dataloader.py:135-151):
 * frames = unique(rows[:,0]); for every window of obs_len + pred_len consecutive frames starting at
 * idx = 0, skip, 2 skip, ... the pedestrians (ascending id) present from its first to its last frame are kept,
 * coordinates rounded to 4 decimals (np.around); a window is kept when it holds MORE than min_ped of them.
 * Outputs (all host): traj_host (cap_peds, obs_len + pred_len, 2) float32 "NTC", non_linear_host (cap_peds)
 * = 1.0 where the quadratic-fit residuals of the last pred_len frames reach `threshold`, peds_in_seq_host
 * (cap_seq) int32; *n_peds_host / *n_seq_host the counts.  traj_host = null only counts.  n_rows bounds both
 * counts.  A pedestrian that spans a window with missing / duplicated frames is ET_ERR_BADARG (ValueError in
 * the reference), as is a frame id that changes under 4-decimal rounding. */
int et_dataset_windows_host(const double* rows_host, int64_t n_rows, int obs_len, int pred_len, int skip,
                            double threshold, int min_ped, float* traj_host, float* non_linear_host,
                            int32_t* peds_in_seq_host, int64_t cap_peds, int64_t cap_seq,
                            int64_t* n_peds_host, int64_t* n_seq_host);

#ifdef __cplusplus
}
#endif
#endif /* ET_B200_H_ */
