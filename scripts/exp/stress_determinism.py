"""Race hunt: the persistent kernels (whole-fit Lloyd, farthest-point seeding, D^2 seeding) repeated many times on shapes
around the one-block-per-SM boundary -- every repetition must reproduce the first one bit for bit, and the whole fit must
equal the launch-per-iteration sequence."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
dev = torch.device("cuda")
bad = 0
for n in (4099, 37_000, 113_665, 113_700, 200_003, 227_329, 500_000, 1_000_000):
    gen = torch.Generator().manual_seed(n)
    data = (torch.randn(1, 6, n, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
    first = n // 3
    c0 = ops.kmeans_farthest_init(data, 20, first)
    u = np.random.RandomState(n).random_sample((1, 20, 4))
    d0 = ops.kmeans_d2_init(data, 20, u)
    acc = ops.KMeansWorkspace(1, 6, 20, dev)
    lab0, cen0 = ops.kmeans_lloyd(data, c0, acc, 25, -1.0)
    km = et.BatchKMeans(n_clusters=20, max_iter=25, tol=-1.0); km.fused = False
    lab_s = km.fit(data, centroids=c0.clone())
    same_step = bool(torch.equal(lab_s, lab0) and torch.equal(km.centroids, cen0))
    fails = {"seed": 0, "d2": 0, "fit": 0}
    for rep in range(60):
        if not torch.equal(ops.kmeans_farthest_init(data, 20, first), c0): fails["seed"] += 1
        if not torch.equal(ops.kmeans_d2_init(data, 20, u), d0): fails["d2"] += 1
        lab, cen = ops.kmeans_lloyd(data, c0, acc, 25, -1.0)
        if not (torch.equal(lab, lab0) and torch.equal(cen, cen0)): fails["fit"] += 1
    bad += sum(fails.values()) + (0 if same_step else 1)
    print(f"n = {n}: whole fit == stepwise {same_step}; non-reproducing repetitions of 60: {fails}", flush=True)
print("STRESS", "FAILED" if bad else "ok")
