import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from oracle import et_oracle as O
l, d, k, n = 1, 8, 32, 4099
gen = torch.Generator().manual_seed(l * 1000 + d * 10 + k)
scale = torch.linspace(4.0, 0.3, d)[None, :, None]
data = (torch.randn(l, d, n, generator=gen) * scale).contiguous().cuda()
cent = data[:, :, torch.randperm(n, generator=gen)[:k].cuda()].contiguous()
for it in (1, 2, 3, 30):
    res = []
    for fused in (True, False):
        km = et.BatchKMeans(n_clusters=k, max_iter=it, tol=-1.0); km.fused = fused
        lab = km.fit(data, centroids=cent.clone()); res.append((lab, km.centroids.clone(), km.n_iter_))
    o_lab, o_cent, o_it, _ = O.kmeans_fit(data.cpu(), k, centroids=cent.cpu().clone(), max_iter=it, tol=-1.0)
    print(it, "fused==stepwise labels", bool(torch.equal(res[0][0], res[1][0])), "cent", bool(torch.equal(res[0][1], res[1][1])),
          "| fused vs oracle mism", int((res[0][0].cpu() != o_lab).sum()), "cent rel", float((res[0][1].cpu() - o_cent).abs().max() / o_cent.abs().max()),
          "| stepwise vs oracle mism", int((res[1][0].cpu() != o_lab).sum()), "cent rel", float((res[1][1].cpu() - o_cent).abs().max() / o_cent.abs().max()))
