import os, sys, torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "scripts"))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
from bench_kernels import time_op
dev = torch.device("cuda"); n = 200_000
o20, p20 = (x.to(dev) for x in synthetic_trajectories(n, seed=1))
hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
dd = et.ETDescriptor(hp).to(dev); dd.parameter_initialization(o20, p20)
Up = dd.U_pred_trunc.detach()
_, _, state = ops.project(o20, p20, dd.U_obs_trunc, Up)
C20 = torch.randn(6, n, 20, device=dev); anchor = torch.randn(6, 20, device=dev)
for _ in range(2):
    a, m = time_op(lambda: ops.reconstruct(C20, Up, state, anchor=anchor), reps=30)
    print(f"reconstruct fwd: avg {1e3*a:.1f} us min {1e3*m:.1f} ({100*n*2428/(a*1e-3)/1e9/6549.1:.1f}%)")
