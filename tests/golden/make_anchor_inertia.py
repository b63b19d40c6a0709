"""usage: PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_anchor_inertia.py eth hotel univ zara1 zara2   (build container, CPU)

Inertia of the farthest-point-seeded BatchKMeans (n_redo = 10: what ETAnchor.anchor_generation runs on the GPU; here the
reference's own CPU BatchKMeans, which the CUDA kernels reproduce) vs sklearn KMeans(n_init=10, random_state=0) as
anchor.py:65-71 calls it, on every scene x {moving, static}."""
import os, sys, json, time
import numpy as np, torch
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference'); sys.path.insert(1, '/root/repo')
os.chdir('/root/reference')
from EigenTrajectory.kmeans import BatchKMeans
from EigenTrajectory.descriptor import ETDescriptor
from utils.dataloader import TrajectoryDataset
from utils.utils import DotDict, augment_trajectory, get_exp_config
from sklearn.cluster import KMeans
torch.set_num_threads(8)
out = {}
for scene in sys.argv[1:]:
    hp = get_exp_config(f"./config/eigentrajectory-{{baseline}}-{scene}.json")
    tr = TrajectoryDataset(f"./datasets/{scene}/train/", obs_len=8, pred_len=12)
    va = TrajectoryDataset(f"./datasets/{scene}/val/", obs_len=8, pred_len=12)
    obs = torch.cat([tr.obs_traj, va.obs_traj]); pred = torch.cat([tr.pred_traj, va.pred_traj])
    obs, pred = augment_trajectory(obs, pred)
    mask = (obs[:, -1] - obs[:, -3]).div(2).norm(p=2, dim=-1) > hp.static_dist
    for tag, m, sca in (("moving", mask, True), ("static", ~mask, False)):
        d = ETDescriptor(hp, norm_sca=sca)
        pred_norm, U = d.parameter_initialization(obs[m], pred[m])
        C = (U.T.detach() @ pred_norm.reshape(-1, 24).T).contiguous()      # (6, N)
        X = C.T.numpy()
        t0 = time.time()
        sk = KMeans(n_clusters=20, random_state=0, init='k-means++', n_init=10).fit(X)
        t_sk = time.time() - t0
        np.random.seed(0)
        km = BatchKMeans(n_clusters=20, n_redo=10)
        t0 = time.time()
        km.fit(C.unsqueeze(0))
        t_km = time.time() - t0
        cent = km.centroids[0].T.numpy()
        ours = float(((X[:, None, :] - cent[None]) ** 2).sum(-1).min(1).sum())
        out[f"{scene}/{tag}"] = dict(n=int(X.shape[0]), sklearn_inertia=float(sk.inertia_), farthest_inertia=ours,
                                     ratio=ours / float(sk.inertia_), t_sklearn=t_sk, t_batchkmeans_cpu=t_km)
        print(scene, tag, out[f"{scene}/{tag}"], flush=True)
json.dump(out, open('/tmp/anchor_eval_%s.json' % '_'.join(sys.argv[1:]), 'w'))     # merged by hand into tests/golden/anchor_inertia.json
