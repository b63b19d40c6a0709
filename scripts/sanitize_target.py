"""compute-sanitizer target: one small-N pass through every CUDA entry point of libet_b200.so (scripts/gpu_sanitize.sh)."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigentrajectory_b200 as et                                   # noqa: E402
from eigentrajectory_b200 import ops                                # noqa: E402
from eigentrajectory_b200.synthetic import synthetic_trajectories   # noqa: E402

N = int(os.environ.get("ET_SANITIZE_N", "4608"))
dev = torch.device("cuda")
hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3, obs_svd=True, pred_svd=True)
obs, pred = (x.to(dev) for x in synthetic_trajectories(N, seed=1))
lib = et.load_library()

# normaliser + descriptor
tn = et.TrajNorm()
tn.calculate_params(obs)
pn = tn.normalize(pred)
tn.denormalize(pn)
d = et.ETDescriptor(hp).to(dev)
d.parameter_initialization(obs, pred)                               # et_gram_init + et_eig_jacobi_pair
d.svd_method = "jacobi"
d.truncated_SVD(pn[:1500])                                          # et_svd_small
d.svd_method = "auto"
d.truncated_SVD(pn)                                                 # et_gram + et_eig_jacobi + et_to_et_space
C_obs, C_pred = d.projection(obs, pred)                             # et_project (TMA projection for N >= 4096)
d.projection(obs[:1000], pred[:1000])                               # plain projection kernel
d.to_Euclidean_space(C_pred, d.U_pred_trunc)
for variant in (1, 2, 3, 4):
    d.project_reconstruct(obs, pred, variant=variant)
d.project_reconstruct(obs[:1001], pred[:1001])
C20 = (torch.randn(6, N, 20, device=dev) * 0.1).requires_grad_(True)
d.projection(obs, pred)
rec = d.reconstruction(C20)                                         # et_reconstruct
rec.square().sum().backward()                                       # et_reconstruct_bwd
anchor = et.ETAnchor(hp).to(dev)
d.reconstruction(C20.detach(), anchor=anchor.C_anchor)

# k-means
data = C_pred.unsqueeze(0).contiguous()
km = et.BatchKMeans(n_clusters=20, max_iter=4)
np.random.seed(0)
cent = km.initialize_centroids(data)                                # persistent seeding kernel
lib.et_tune(5, 1)
km.initialize_centroids(data)                                       # per-step seeding kernels
lib.et_tune(5, 0)
ms, lab = km.get_labels(data, cent)                                 # et_kmeans_assign
km.compute_centroids(data, lab)                                     # et_kmeans_accumulate + et_kmeans_finalize
km.fit(data, centroids=cent)                                        # et_kmeans_lloyd (persistent, grid barriers)
km.fused = False
km.fit(data, centroids=cent)                                        # assign / finalize launch pairs
km3 = et.BatchKMeans(n_clusters=5, max_iter=3)
km3.fit(torch.randn(3, 4, 700, device=dev))                         # generic (d, K) instantiation, batch of 3
ops.kmeans_seed_step(data, cent, 7)
key = ops.kmeans_seed_candidate(data, cent, 7, 100)
ops.kmeans_seed_fetch(data, 100, key)

# metrics
gt = pred
ops.ade_fde(rec.detach(), gt, want_argmin=True, want_tcc=True)
ops.col(rec.detach()[:, :60].contiguous())

# fused forward + losses + backward through the hook seam
W = torch.nn.Parameter(torch.randn(120, 8, device=dev) * 0.05)


class Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.W = W

    def forward(self, x):
        return (self.W @ x).reshape(6, 20, -1).permute(0, 2, 1)


hook = types.SimpleNamespace(model_forward_pre_hook=lambda C, o, info=None: torch.cat([C, o], dim=0),
                             model_forward=lambda x, m: m(x), model_forward_post_hook=lambda y, info=None: y)
model = et.EigenTrajectory(Stub(), hook, hp).to(dev)
model.calculate_parameters(obs, pred)
out = model(obs[:57], pred[:57])
(out["loss_eigentraj"] + out["loss_euclidean_ade"] + out["loss_euclidean_fde"] + out["recon_traj"].mean()).backward()
model.fused = False
model(obs[:57], pred[:57])
torch.cuda.synchronize()
print(f"sanitize target done: {et.launch_count()} launches")
