"""Times et_eig_jacobi (24x24 and 16x16 Gram matrices of 1e6 synthetic pedestrians) for every block-size variant."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_000_000, seed=0))
Go, Gp = ops.gram(obs, pred, True, True, True)
lib = et.load_library()
ref = {}
for name, G, variants in (("24x24", Gp, (32, 96, 144, 288)), ("16x16", Go, (32, 64, 128))):
    for nt in variants:
        lib.et_tune(3, nt)
        info = torch.zeros(2, dtype=torch.int32, device=dev)
        U, S, U64, S64 = ops.eig_basis(G, 6, want64=True, info=info)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        evs[0].record()
        for i in range(10):
            ops.eig_basis(G, 6)
            evs[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(10))
        key = (name,)
        same = "ref" if key not in ref else f"U max diff vs 32-thread {float((U64 - ref[key][0]).abs().max()):.2e} S rel {float(((S64 - ref[key][1]) / ref[key][1]).abs().max()):.2e}"
        ref.setdefault(key, (U64, S64))
        print(f"{name} threads {nt:3d}: median {1e3 * ts[5]:7.1f} us  min {1e3 * ts[0]:7.1f} us  sweeps/rot {info.tolist()}  {same}")
lib.et_tune(3, 0)
