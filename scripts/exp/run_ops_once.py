"""Runs every hot-path op a few times (for ncu captures)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
n, n20 = 1_000_000, 200_000
hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
obs, pred = (x.to(dev) for x in synthetic_trajectories(n, seed=0))
desc = et.ETDescriptor(hp).to(dev)
for rep in range(3):
    desc.parameter_initialization(obs, pred)
    Uo, Up = desc.U_obs_trunc.detach(), desc.U_pred_trunc.detach()
    ops.project_reconstruct(obs, pred, Uo, Up)
    C_obs, C_pred, state = ops.project(obs[:n20].contiguous(), pred[:n20].contiguous(), Uo, Up)
    C20 = torch.randn(6, n20, 20, device=dev).requires_grad_(True)
    rec = ops.reconstruct(C20, Up, state, anchor=torch.randn(6, 20, device=dev))
    rec.sum().backward()
    ops.ade_fde(rec.detach(), pred[:n20].contiguous())
    data = ops.project(obs, pred, Uo, Up)[1].unsqueeze(0).contiguous()
    km = et.BatchKMeans(n_clusters=20, max_iter=3)
    np.random.seed(0)
    km.fit(data)
    small = obs[:181].contiguous()
    st = ops.norm_params(small, True, True, False)
    ops.svd_small(ops.normalize(small, *st), 6)
torch.cuda.synchronize()
print("done")
