#!/bin/bash
mkdir -p gpurun_out
# (tests run by the calling script)
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
dev = torch.device("cuda")
gen = torch.Generator().manual_seed(1234)
data = (torch.randn(1, 6, 1_000_000, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
km = et.BatchKMeans(n_clusters=20); np.random.seed(0)
cent = km.initialize_centroids(data)
acc = ops.KMeansWorkspace(1, 6, 20, dev)
for _ in range(2): ops.kmeans_lloyd(data, cent, acc, 100, -1.0, want_labels=False)
torch.cuda.synchronize()
ts = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.kmeans_lloyd(data, cent, acc, 100, -1.0, want_labels=False); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 10)
print("whole-fit kernel: us per Lloyd iteration", [round(t, 2) for t in ts])
PY
