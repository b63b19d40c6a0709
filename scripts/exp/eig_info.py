import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_000_000, seed=0))
for flags in ((True, True, True), (True, True, False)):
    Go, Gp = ops.gram(obs, pred, *flags)
    for G in (Go, Gp):
        info = torch.zeros(2, dtype=torch.int32, device=dev)
        U, S = ops.eig_basis(G, 6, info=info)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.eig_basis(G, 6, info=info); e1.record(); torch.cuda.synchronize()
        print("m", G.size(0), "flags", flags, "sweeps/rotations", info.tolist(), "ms", e0.elapsed_time(e1), "S", S.tolist()[:3])
