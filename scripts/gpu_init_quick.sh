#!/bin/bash
# init path timing: Gram pass, eigen-solves (two-barrier body vs the four-barrier one), parameter_initialization
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
lib = et.load_library(); dev = torch.device("cuda")
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_000_000, seed=0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=20, cold=False):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        if cold: flush.fill_(1); flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return round(ts[len(ts) // 2], 1)
G_o, G_p, _, _ = ops.gram_init(obs, pred)
info = torch.zeros(2, dtype=torch.int32, device=dev)
print("gram (both matrices, normalise fused) us:", timed(lambda: ops.gram(obs, pred, True, True, True), cold=True))
print("gram_init (+ state + normalised futures) us:", timed(lambda: ops.gram_init(obs, pred), cold=True))
for tag, knob in (("two-barrier body, second generation + symmetric update (default)", 0), ("two-barrier body, second generation", 2002), ("two-barrier body, first generation", 2001), ("four-barrier body 288/128 threads", 288)):
    lib.et_tune(3, knob)
    print(f"eig 24x24 {tag} us:", timed(lambda: ops.eig_basis(G_p, 6, info=info)), "sweeps/rotations", info.tolist())
    print(f"eig 16x16 {tag} us:", timed(lambda: ops.eig_basis(G_o, 6, info=info)), "sweeps/rotations", info.tolist())
for tag, knob in (("first generation, 1 row per V thread", 1001), ("first generation, 1 Newton step per rsqrt", 1003)):
    lib.et_tune(3, knob)
    U, S, U64, S64 = ops.eig_basis(G_p, 6, want64=True)
    lib.et_tune(3, 0)
    U0, S0, U640, S640 = ops.eig_basis(G_p, 6, want64=True)
    lib.et_tune(3, knob)
    print(f"eig 24x24 two-barrier body, {tag} us:", timed(lambda: ops.eig_basis(G_p, 6, info=info)), "sweeps/rotations", info.tolist(),
          "projector diff vs default %.1e, S rel %.1e" % (float((U64 @ U64.T - U640 @ U640.T).norm()), float(((S64 - S640).abs() / S640).max())))
lib.et_tune(3, 0)
print("eig pair (both bases, one launch) us:", timed(lambda: ops.eig_basis_pair(G_o, G_p, 6)))
d = et.ETDescriptor(et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)).to(dev)
print("ETDescriptor.parameter_initialization (N = 1e6) us:", timed(lambda: d.parameter_initialization(obs, pred), cold=True))
PY
