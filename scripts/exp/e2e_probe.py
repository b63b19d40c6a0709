"""Experiment: where does the host-buffer path spend its time?  PCIe bandwidths, pinned allocation cost, pipeline."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories

dev = torch.device("cuda")
n = 1_000_000
obs, pred = synthetic_trajectories(n, seed=0)
t0 = time.perf_counter(); obs_p, pred_p = obs.pin_memory(), pred.pin_memory(); print("pin inputs %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for _ in range(3):
    t0 = time.perf_counter(); x = torch.empty((n, 12, 2), pin_memory=True); print("alloc pinned 96MB %.2f ms" % (1e3 * (time.perf_counter() - t0))); del x

def bw(fn, nbytes, name, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {dt*1e3:.2f} ms  {nbytes/dt/1e9:.1f} GB/s")

d_pred = torch.empty((n, 12, 2), device=dev); h_out = torch.empty((n, 12, 2), pin_memory=True)
bw(lambda: d_pred.copy_(pred_p, non_blocking=True), n * 96, "H2D pinned 96MB")
bw(lambda: h_out.copy_(d_pred, non_blocking=True), n * 96, "D2H pinned 96MB")
bw(lambda: d_pred.copy_(pred, non_blocking=True), n * 96, "H2D pageable 96MB")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_pred.copy_(pred_p, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_pred, non_blocking=True)
    s1.synchronize(); s2.synchronize()
bw(both, n * 192, "H2D + D2H concurrently (192MB)")

hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
desc = et.ETDescriptor(hp).to(dev)
desc.parameter_initialization(obs.to(dev), pred.to(dev))
for chunk in (65536, 131072, 262144):
    ops.HOST_CHUNK = chunk
    for wc in (True, False):
        ts = []
        for i in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = desc.project_reconstruct(obs_p, pred_p, want_coeffs=wc)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"chunk {chunk} coeffs={wc}: " + " ".join(f"{1e3*t:.1f}" for t in ts) + " ms")
