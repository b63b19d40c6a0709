"""Sweep of the host-buffer pipeline of ETDescriptor.project_reconstruct: chunk size x stream count (one GPU)."""
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
n = 1_000_000
obs, pred = synthetic_trajectories(n, seed=0)
ho, hp_ = obs.pin_memory(), pred.pin_memory()
d = et.ETDescriptor(et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)).to(dev)
d.parameter_initialization(obs.to(dev), pred.to(dev))
def run(reps=8):
    ts = []
    out = None
    for j in range(reps + 3):
        del out
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = d.project_reconstruct(ho, hp_)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return 1e3 * sum(ts[3:]) / reps
for chunk in (65536, 131072, 196608, 262144):
    for first in (16384, 32768):
        for streams in (3, 4, 6):
            ops.HOST_CHUNK, ops.HOST_FIRST_CHUNK = chunk, first
            ops._side_streams.clear()
            orig = ops._streams
            ops._streams = (lambda device, count=streams, _o=orig: _o(device, count))
            ms = run()
            ops._streams = orig
            print(json.dumps({"chunk": chunk, "first": first, "streams": streams, "ms_per_step": round(ms, 3), "M_traj_per_s": round(n / ms / 1e3, 1)}), flush=True)
