import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "scripts"))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from bench_kernels import time_op
dev = torch.device("cuda")
lib = et.load_library()
n = 200_000
gt = torch.randn(n, 12, 2, device=dev).cumsum(1)
pred = gt[None] + torch.randn(20, n, 12, 2, device=dev) * 0.4
for cfg in (13, 14, 15, 16, 17, 13):
    lib.et_tune(0, cfg)
    a, m = time_op(lambda: ops.ade_fde(pred, gt), reps=30)
    print(f"ADE config {cfg}: avg {1e3*a:.1f} us min {1e3*m:.1f} us -> {n*2024/(a*1e-3)/1e9:.0f} GB/s ({100*n*2024/(a*1e-3)/1e9/6549.1:.1f}%)")
lib.et_tune(0, 0)
# projection, new TMA pipeline
from eigentrajectory_b200.synthetic import synthetic_trajectories
N = 1_000_000
obs, pred2 = (x.to(dev) for x in synthetic_trajectories(N, seed=0))
hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
d = et.ETDescriptor(hp).to(dev); d.parameter_initialization(obs, pred2)
a, m = time_op(lambda: ops.project(obs, pred2, d.U_obs_trunc, d.U_pred_trunc), reps=30)
print(f"project (TMA pipeline): avg {1e3*a:.1f} us -> {N*236/(a*1e-3)/1e9:.0f} GB/s ({100*N*236/(a*1e-3)/1e9/6549.1:.1f}%)")
a, m = time_op(lambda: ops.project(obs[:999_999], pred2[:999_999], d.U_obs_trunc, d.U_pred_trunc), reps=30)
print(f"project (thread-per-row, N%4!=0): avg {1e3*a:.1f} us -> {N*236/(a*1e-3)/1e9:.0f} GB/s")
