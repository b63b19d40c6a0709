import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops, parallel as P
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_250_000, seed=1000))
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print(json.dumps({"basis_ms_1gpu_1.25e6_rows": timed(lambda: P.sharded_basis(obs, pred, 6))}))
