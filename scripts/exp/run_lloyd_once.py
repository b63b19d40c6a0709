"""Runs the whole-fit k-means kernel a few times (for ncu captures): 20 Lloyd iterations per launch, N = 1e6."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
dev = torch.device("cuda")
n = int(os.environ.get("N", 1_000_000))
gen = torch.Generator().manual_seed(1234)
data = (torch.randn(1, 6, n, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
km = et.BatchKMeans(n_clusters=20)
np.random.seed(0)
cent = km.initialize_centroids(data)
acc = ops.KMeansWorkspace(1, 6, 20, dev)
for rep in range(3):
    ops.kmeans_lloyd(data, cent, acc, int(os.environ.get("ITERS", 20)), -1.0, want_labels=False)
torch.cuda.synchronize()
print("done", acc.status.tolist())
