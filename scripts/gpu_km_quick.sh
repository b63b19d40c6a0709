#!/bin/bash
# k-means timing: whole-fit kernel per Lloyd iteration (tournament vs serial arg-max), seeding (persistent vs per step)
mkdir -p gpurun_out
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
lib = et.load_library()
dev = torch.device("cuda")
gen = torch.Generator().manual_seed(1234)
data = (torch.randn(1, 6, 1_000_000, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
km = et.BatchKMeans(n_clusters=20); np.random.seed(0)
cent = km.initialize_centroids(data)
acc = ops.KMeansWorkspace(1, 6, 20, dev)

def per_iter(iters=100, reps=7):
    for _ in range(2): ops.kmeans_lloyd(data, cent, acc, iters, -1.0, want_labels=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.kmeans_lloyd(data, cent, acc, iters, -1.0, want_labels=False); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / iters)
    return [round(t, 2) for t in ts]

print("whole-fit kernel, 12 warps, ascending arg-max scan (default): us per Lloyd iteration", per_iter())
lib.et_tune(6, 1)
print("whole-fit kernel, 12 warps, per-group tournament arg-max:     us per Lloyd iteration", per_iter())
lib.et_tune(6, 0)
ref_labels, ref_cent = ops.kmeans_lloyd(data, cent, acc, 30, -1.0)
for w in (16,):
    lib.et_tune(7, w)
    print(f"whole-fit kernel, {w} warps on shared record columns: us per Lloyd iteration", per_iter())
    lab, c = ops.kmeans_lloyd(data, cent, acc, 30, -1.0)
    print(f"   vs 12 warps after 30 iterations: label mismatches {int((lab != ref_labels).sum())}, centroids rel "
          f"{float((c - ref_cent).abs().max() / ref_cent.abs().max()):.2e}")
lib.et_tune(7, 0)

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / reps, 1)

print("assign only (labels + maxsims written) us:", timed(lambda: ops.kmeans_assign(data, cent)))
print("seeding persistent us:", timed(lambda: ops.kmeans_farthest_init(data, 20, 12345)))
lib.et_tune(5, 1)
print("seeding per-step launches us:", timed(lambda: ops.kmeans_farthest_init(data, 20, 12345)))
lib.et_tune(5, 0)
km = et.BatchKMeans(n_clusters=20, max_iter=100, tol=-1.0)
np.random.seed(0)
print("BatchKMeans.fit (seeding + 100 iterations + labels) ms:", timed(lambda: km.fit(data), 5) / 1e3)
PY
