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
from eigentrajectory_b200.synthetic import synthetic_trajectories as syn
o20, p20 = (x.to(dev) for x in syn(n, seed=1))
hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
dd = et.ETDescriptor(hp).to(dev); dd.parameter_initialization(o20, p20)
Up = dd.U_pred_trunc.detach()
_, _, state = ops.project(o20, p20, dd.U_obs_trunc, Up)
C20 = torch.randn(6, n, 20, device=dev)
anchor = torch.randn(6, 20, device=dev)
from eigentrajectory_b200._lib import ptr, stream_of, check
rec = ops.reconstruct(C20, Up, state, anchor=anchor); gC = torch.empty_like(C20)
def bwd():
    check(lib.et_reconstruct_bwd(ptr(rec), n, 20, 6, 12, ptr(Up), 7, ptr(state[1]), ptr(state[2]), ptr(gC), stream_of(dev)), "bwd")
for knob in (0, -1, 4, 3, 2, 0):
    lib.et_tune(1, knob)
    a, m = time_op(lambda: ops.reconstruct(C20, Up, state, anchor=anchor), reps=30)
    a2, m2 = time_op(bwd, reps=30)
    print(f"REC knob {knob}: fwd avg {1e3*a:.1f} us ({100*n*2428/(a*1e-3)/1e9/6549.1:.1f}%)   bwd avg {1e3*a2:.1f} us ({100*n*2420/(a2*1e-3)/1e9/6549.1:.1f}%)")
lib.et_tune(1, 0)
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
