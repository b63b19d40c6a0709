"""ncu target: one whole-fit launch (10 Lloyd iterations, 1e6 points, K = 20) after two warm-up launches."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
dev = torch.device("cuda")
gen = torch.Generator().manual_seed(1234)
data = (torch.randn(1, 6, 1_000_000, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
km = et.BatchKMeans(n_clusters=20); np.random.seed(0)
cent = km.initialize_centroids(data)
acc = ops.KMeansWorkspace(1, 6, 20, dev)
for _ in range(3):
    ops.kmeans_lloyd(data, cent, acc, int(os.environ.get("KM_ITERS", "10")), -1.0, want_labels=False)
torch.cuda.synchronize()
