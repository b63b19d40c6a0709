"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz.

Run once in the build container (CPU, torch 2.11, sklearn 1.9):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference from /root/reference (read-only) and pushes committed,
seeded inputs through the reference's own classes.  The .npz files carry both the
inputs and the reference outputs, so tests on the GPU box (where /root/reference
does not exist) never need the reference or this script.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ET_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from EigenTrajectory import EigenTrajectory, TrajNorm            # noqa: E402  (reference)
from EigenTrajectory.descriptor import ETDescriptor              # noqa: E402
from EigenTrajectory.anchor import ETAnchor                      # noqa: E402
from EigenTrajectory.kmeans import BatchKMeans                   # noqa: E402
from utils.metrics import compute_batch_ade, compute_batch_fde, compute_batch_tcc, compute_batch_col   # noqa: E402
from utils.utils import DotDict, augment_trajectory              # noqa: E402
from utils.dataloader import TrajectoryDataset                   # noqa: E402

from oracle.et_oracle import synthetic_trajectories              # noqa: E402  (input generator only)

torch.set_num_threads(8)
HP = DotDict(dict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.419,
                  obs_svd=True, pred_svd=True))


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB, keys={sorted(arrs)}")


def eth_test():
    """Config 1: ETH test split through script/descriptor_evaluation.py:22-36,87-112."""
    cwd = os.getcwd()
    os.chdir(REF)
    ds = TrajectoryDataset("./datasets/eth/test/", obs_len=8, pred_len=12)
    os.chdir(cwd)
    obs, pred = ds.obs_traj, ds.pred_traj
    n = obs.shape[0]
    tn = TrajNorm(ori=True, rot=True, sca=False)
    tn.calculate_params(obs)
    on, pn = tn.normalize(obs), tn.normalize(pred)
    A, B = on.reshape(n, 16).T, pn.reshape(n, 24).T
    Uo, So, _ = torch.linalg.svd(A, full_matrices=False)
    Up, Sp, _ = torch.linalg.svd(B, full_matrices=False)
    eo, ep = [], []
    for k in range(1, 13):
        Ar = Uo[:, :k] @ (Uo[:, :k].T @ A)
        Br = Up[:, :k] @ (Up[:, :k].T @ B)
        ro = tn.denormalize(Ar.T.reshape(n, 8, 2))
        rp = tn.denormalize(Br.T.reshape(n, 12, 2))
        eo.append((ro - obs).norm(p=2, dim=-1).mean().item())
        ep.append((rp - pred).norm(p=2, dim=-1).mean().item())
    save("eth_test.npz", obs=npy(obs), pred=npy(pred), U_obs=npy(Uo), S_obs=npy(So), U_pred=npy(Up),
         S_pred=npy(Sp), err_obs=np.array(eo), err_pred=np.array(ep),
         obs_norm=npy(on), pred_norm=npy(pn), rot=npy(tn.traj_rot),
         num_peds_in_seq=np.asarray(ds.num_peds_in_seq))


def eth_init():
    """Init path on ETH train+val (flip-augmented): the four singular spectra + bases."""
    cwd = os.getcwd()
    os.chdir(REF)
    tr = TrajectoryDataset("./datasets/eth/train/", obs_len=8, pred_len=12)
    va = TrajectoryDataset("./datasets/eth/val/", obs_len=8, pred_len=12)
    os.chdir(cwd)
    obs = torch.cat([tr.obs_traj, va.obs_traj], 0)
    pred = torch.cat([tr.pred_traj, va.pred_traj], 0)
    # raw (un-augmented) inputs, so the GPU tests can rebuild the init set without the dataset files
    save("eth_init_data.npz", obs=npy(obs), pred=npy(pred))
    obs, pred = augment_trajectory(obs, pred)
    mask = (obs[:, -1] - obs[:, -3]).div(2).norm(p=2, dim=-1) > HP.static_dist
    out = dict(n_total=np.array(obs.shape[0]), n_moving=np.array(int(mask.sum())))
    for tag, m, sca in (("m", mask, True), ("s", ~mask, False)):
        d = ETDescriptor(HP, norm_sca=sca)
        on, pn = d.normalize_trajectory(obs[m], pred[m])
        Uo, So, _ = d.truncated_SVD(on)
        Up, Sp, _ = d.truncated_SVD(pn)
        # fp64 truth on the same normalised data
        Uo64, So64, _ = torch.linalg.svd(on.double().reshape(-1, 16).T, full_matrices=False)
        Up64, Sp64, _ = torch.linalg.svd(pn.double().reshape(-1, 24).T, full_matrices=False)
        out.update({f"U_obs_{tag}": npy(Uo), f"S_obs_{tag}": npy(So), f"U_pred_{tag}": npy(Up),
                    f"S_pred_{tag}": npy(Sp), f"U_obs64_{tag}": npy(Uo64[:, :6]), f"S_obs64_{tag}": npy(So64[:6]),
                    f"U_pred64_{tag}": npy(Up64[:, :6]), f"S_pred64_{tag}": npy(Sp64[:6])})
    save("eth_init.npz", **out)


def descriptor_syn():
    """Configs 2-shaped: seeded synthetic N=1536 through ETDescriptor (sca on / off)."""
    obs, pred = synthetic_trajectories(1536, seed=0)
    out = dict(obs=npy(obs), pred=npy(pred))
    g = torch.Generator().manual_seed(7)
    n_rec = 100
    C_in = torch.randn(6, n_rec, 20, generator=g) * torch.tensor([8., 3., 1., .5, .3, .2])[:, None, None]
    out["C_in"] = npy(C_in)
    for tag, sca in (("sca1", True), ("sca0", False)):
        d = ETDescriptor(HP, norm_sca=sca)
        pred_norm, U_pred = d.parameter_initialization(obs, pred)
        on, _ = d.normalize_trajectory(obs, pred)
        _, So, Vo = d.truncated_SVD(on)
        _, Sp, Vp = d.truncated_SVD(pred_norm)
        C_obs, C_pred = d.projection(obs, pred)
        tn = d.traj_normalizer
        out.update({f"U_obs_{tag}": npy(d.U_obs_trunc), f"U_pred_{tag}": npy(d.U_pred_trunc),
                    f"S_obs_{tag}": npy(So), f"S_pred_{tag}": npy(Sp),
                    f"V_obs_{tag}": npy(Vo), f"V_pred_{tag}": npy(Vp),
                    f"obs_norm_{tag}": npy(on), f"pred_norm_{tag}": npy(pred_norm),
                    f"C_obs_{tag}": npy(C_obs), f"C_pred_{tag}": npy(C_pred),
                    f"ori_{tag}": npy(tn.traj_ori), f"rot_{tag}": npy(tn.traj_rot)})
        if sca:
            out[f"sca_{tag}"] = npy(tn.traj_sca)
        # rank-6 round trip of the whole set (S=1 shape of the headline op)
        rec_obs = d.denormalize_trajectory(d.to_Euclidean_space(C_obs, d.U_obs_trunc))
        rec_pred = d.denormalize_trajectory(d.to_Euclidean_space(C_pred, d.U_pred_trunc))
        out[f"rec_obs_{tag}"], out[f"rec_pred_{tag}"] = npy(rec_obs), npy(rec_pred)
        # S=20 reconstruction on the first n_rec pedestrians (state = that sub-batch)
        d.projection(obs[:n_rec], pred[:n_rec])
        rec = d.reconstruction(C_in)
        out[f"recon20_{tag}"] = npy(rec)
        # anchor add + reconstruction gradient wrt C
        a = ETAnchor(HP)
        a.C_anchor.data = torch.randn(6, 20, generator=g)
        Cg = C_in.clone().requires_grad_(True)
        rec2 = d.reconstruction(a(Cg))
        w = torch.randn(rec2.shape, generator=g)
        (rec2 * w).sum().backward()
        out[f"anchor_{tag}"], out[f"recon20_anchor_{tag}"] = npy(a.C_anchor), npy(rec2)
        out[f"grad_w_{tag}"], out[f"grad_C_{tag}"] = npy(w), npy(Cg.grad)
    save("descriptor_syn.npz", **out)


def kmeans():
    """Config 3-shaped: BatchKMeans on (2,6,4096) scale-decay Gaussians, lock-step trace."""
    g = torch.Generator().manual_seed(1234)
    scale = torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]
    data = (torch.randn(2, 6, 4096, generator=g) * scale).contiguous()
    km = BatchKMeans(n_clusters=20)
    np.random.seed(0)
    first = np.random.randint(4096)
    np.random.seed(0)
    c0 = km.initialize_centroids(data)
    labels_t, maxsims_t, cents_t = [], [], [c0.clone()]
    c = c0
    for _ in range(12):
        ms, lb = km.get_labels(data, c)
        c = km.compute_centroids(data, lb)
        labels_t.append(npy(lb).astype(np.int8))
        maxsims_t.append(npy(ms))
        cents_t.append(c.clone())
    np.random.seed(0)
    km2 = BatchKMeans(n_clusters=20)
    fit_labels = km2.fit(data)
    # ragged / edge shapes for the assign step alone
    edge = {}
    for n in (1, 31, 33, 1000):
        dd = (torch.randn(1, 6, n, generator=g) * scale).contiguous()
        cc = (torch.randn(1, 6, 20, generator=g) * scale).contiguous()
        ms, lb = km.get_labels(dd, cc)
        edge[f"edge{n}_data"], edge[f"edge{n}_cent"] = npy(dd), npy(cc)
        edge[f"edge{n}_labels"], edge[f"edge{n}_maxsims"] = npy(lb), npy(ms)
    save("kmeans.npz", data=npy(data), first_index=np.array(first), init_centroids=npy(c0),
         trace_labels=np.stack(labels_t), trace_maxsims=np.stack(maxsims_t),
         trace_centroids=np.stack([npy(x) for x in cents_t]),
         fit_labels=npy(fit_labels).astype(np.int8), fit_centroids=npy(km2.centroids), **edge)


def metrics():
    g = torch.Generator().manual_seed(99)
    gt = torch.randn(300, 12, 2, generator=g).cumsum(1)
    pred = gt[None] + torch.randn(20, 300, 12, 2, generator=g) * 0.4
    # a crowded scene for the collision metric: 48 pedestrians walking through a 6 m x 6 m area
    p0 = torch.rand(48, 1, 2, generator=g) * 6
    v = torch.randn(48, 1, 2, generator=g) * 0.4
    scene_gt = p0 + v * torch.arange(12)[None, :, None]
    scene_pred = scene_gt[None] + torch.randn(20, 48, 12, 2, generator=g) * 0.15
    save("metrics.npz", pred=npy(pred), gt=npy(gt), ade=compute_batch_ade(pred, gt),
         fde=compute_batch_fde(pred, gt[None]), tcc=compute_batch_tcc(pred, gt), col=compute_batch_col(pred, gt),
         scene_pred=npy(scene_pred), scene_gt=npy(scene_gt), scene_col=compute_batch_col(scene_pred, scene_gt),
         scene_tcc=compute_batch_tcc(scene_pred, scene_gt[None]))


def model_forward():
    """EigenTrajectory.forward on CPU through the hook seam with a fixed linear stub baseline."""
    g = torch.Generator().manual_seed(5)
    W = torch.randn(6 * 20, 8, generator=g) * 0.1

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.W = torch.nn.Parameter(W.clone())

        def forward(self, x):                     # x (8,N) -> (6,N,20)
            return (self.W @ x).reshape(6, 20, -1).permute(0, 2, 1)

    hook = types.SimpleNamespace(
        model_forward_pre_hook=lambda C, o, info=None: torch.cat([C, o], dim=0),
        model_forward=lambda x, m: m(x),
        model_forward_post_hook=lambda y, info=None: y)
    stub = Stub()
    model = EigenTrajectory(stub, hook, HP)
    obs, pred = synthetic_trajectories(3000, seed=3)
    # make ~1/3 of pedestrians slow so both the moving and the static groups are populated
    slow = torch.arange(3000) % 3 == 0
    c = obs[:, -1:, :].clone()
    obs = torch.where(slow[:, None, None], c + (obs - c) * 0.3, obs)
    pred = torch.where(slow[:, None, None], c + (pred - c) * 0.3, pred)
    model.calculate_parameters(obs, pred)
    sd = {k: npy(v) for k, v in model.state_dict().items() if k.startswith("ET_")}
    o, p = obs[:57].clone(), pred[:57].clone()
    out = model(o, p)
    loss = out["loss_eigentraj"] + out["loss_euclidean_ade"] + out["loss_euclidean_fde"]
    loss.backward()
    test_out = model(o)
    save("model_forward.npz", W=npy(W), obs=npy(o), pred=npy(p), init_obs=npy(obs), init_pred=npy(pred),
         recon=npy(out["recon_traj"]), loss_eigentraj=npy(out["loss_eigentraj"]),
         loss_ade=npy(out["loss_euclidean_ade"]), loss_fde=npy(out["loss_euclidean_fde"]),
         grad_W=npy(stub.W.grad), recon_test=npy(test_out["recon_traj"]),
         **{"sd_" + k: v for k, v in sd.items()})


def _synthetic_scene_text(seed, n_frames, n_peds, frame_step=10, drop_frames=(), decimals=6):
    """A dataset file in the reference's format: pedestrians with contiguous random lifetimes, a few frames
    missing altogether, coordinates with more than 4 decimals (exercises np.around)."""
    rng = np.random.RandomState(seed)
    frames = [f * frame_step for f in range(n_frames) if f not in drop_frames]
    rows = []
    for ped in range(1, n_peds + 1):
        start = rng.randint(0, len(frames) - 2)
        life = rng.randint(1, 45)
        p = rng.uniform(-8, 8, size=2)
        v = rng.uniform(-0.6, 0.6, size=2)
        curve = rng.normal(0, 0.02, size=2) if ped % 3 == 0 else np.zeros(2)
        for j, fi in enumerate(range(start, min(start + life, len(frames)))):
            wiggle = 0.12 * np.sin(0.9 * j + ped) if ped % 4 == 0 else 0.0      # not a quadratic: non-linear flag
            pos = p + v * j + curve * j * j + wiggle + rng.normal(0, 0.002, size=2)
            rows.append((frames[fi], float(ped), pos[0], pos[1]))
    rows.sort(key=lambda r: (r[0], rng.rand()))          # frame-major, pedestrians in arbitrary order inside a frame
    return "".join(f"{fr:d}\t{ped:.1f}\t{x:.{decimals}f}\t{y:.{decimals}f}\n" for fr, ped, x, y in rows)


def dataset():
    """utils/dataloader.py: TrajectoryDataset / TrajBatchSampler / traj_collate_fn on the ETH test file and on
    synthetic two-file splits (text inputs are stored with the outputs)."""
    import tempfile
    from utils.dataloader import TrajBatchSampler, traj_collate_fn

    out = {}

    def run(tag, files, **kw):
        with tempfile.TemporaryDirectory() as tmp:
            for name, text in files.items():
                with open(os.path.join(tmp, name), "w") as f:
                    f.write(text)
            ds = TrajectoryDataset(tmp + "/", **kw)
            order = os.listdir(tmp)
        out[f"{tag}_file_order"] = np.array(order)
        for name, text in files.items():
            out[f"{tag}_text_{name}"] = np.frombuffer(text.encode(), dtype=np.uint8)
        out[f"{tag}_obs"] = npy(ds.obs_traj.contiguous())
        out[f"{tag}_pred"] = npy(ds.pred_traj.contiguous())
        out[f"{tag}_loss_mask"] = npy(ds.loss_mask)
        out[f"{tag}_non_linear"] = npy(ds.non_linear_ped)
        out[f"{tag}_num_peds_in_seq"] = np.asarray(ds.num_peds_in_seq)
        out[f"{tag}_seq_start_end"] = np.asarray(ds.seq_start_end)
        return ds

    with open(os.path.join(REF, "datasets/eth/test/biwi_eth.txt")) as f:
        eth_text = f.read()
    ds = run("eth", {"biwi_eth.txt": eth_text}, obs_len=8, pred_len=12)
    # test-phase batching (no shuffle): batch index lists and one collated batch
    batches = list(TrajBatchSampler(ds, batch_size=32, shuffle=False, drop_last=False))
    out["eth_batch_sizes"] = np.array([len(b) for b in batches])
    out["eth_batch_first"] = np.array(batches[0])
    col = traj_collate_fn([ds[i] for i in batches[3]])
    out["eth_b3_obs"], out["eth_b3_pred"] = npy(col[0]), npy(col[1])
    out["eth_b3_scene_mask"], out["eth_b3_seq_start_end"] = npy(col[4]), npy(col[5])

    syn = {"a.txt": _synthetic_scene_text(1, 90, 60, drop_frames=(17, 40)),
           "b.txt": _synthetic_scene_text(2, 70, 45, frame_step=6, decimals=5)}
    run("syn", syn, obs_len=8, pred_len=12)
    run("syn_skip3", syn, obs_len=8, pred_len=12, skip=3)
    run("syn_short", syn, obs_len=3, pred_len=5, min_ped=0, threshold=0.002)
    save("dataset.npz", **out)


def eth_eval():
    """The evaluation loop of ETTrainer.test (utils/trainer.py:172-195) on the ETH test split, on CPU: reference
    dataloader (batch_size=1) -> EigenTrajectory.forward through the hook seam (stub predictor, bases / anchors of
    model_forward.npz) -> ADE / FDE / TCC / COL per pedestrian -> AverageMeter means."""
    import tempfile
    from utils.dataloader import get_dataloader
    from utils.metrics import AverageMeter

    mf = np.load(os.path.join(HERE, "model_forward.npz"))
    ds = np.load(os.path.join(HERE, "dataset.npz"))
    W = torch.from_numpy(mf["W"])

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.W = torch.nn.Parameter(W.clone())

        def forward(self, x):
            return (self.W @ x).reshape(6, 20, -1).permute(0, 2, 1)

    hook = types.SimpleNamespace(
        model_forward_pre_hook=lambda C, o, info=None: torch.cat([C, o], dim=0),
        model_forward=lambda x, m: m(x),
        model_forward_post_hook=lambda y, info=None: y)
    model = EigenTrajectory(Stub(), hook, HP)
    sd = {k[3:]: torch.from_numpy(mf[k]) for k in mf.files if k.startswith("sd_")}
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    model.eval()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "test"))
        with open(os.path.join(tmp, "test", "biwi_eth.txt"), "wb") as f:
            f.write(ds["eth_text_biwi_eth.txt"].tobytes())
        loader = get_dataloader(tmp, "test", 8, 12, batch_size=1)
        funcs = {"ADE": compute_batch_ade, "FDE": compute_batch_fde, "TCC": compute_batch_tcc, "COL": compute_batch_col}
        meters = {k: AverageMeter() for k in funcs}
        with torch.no_grad():
            for batch in loader:
                obs, pred = batch[:2]
                out = model(obs)
                for k, fn in funcs.items():
                    meters[k].extend(fn(out["recon_traj"], pred))
    arrs = {f"per_ped_{k}": np.concatenate(m.data, axis=0) for k, m in meters.items()}
    arrs.update({f"mean_{k}": np.float64(m.mean()) for k, m in meters.items()})
    print({k: float(v) for k, v in arrs.items() if k.startswith("mean_")})
    save("eth_eval.npz", **arrs)


if __name__ == "__main__":
    which = sys.argv[1:] or ["eth_test", "eth_init", "descriptor_syn", "kmeans", "metrics", "model_forward", "dataset", "eth_eval"]
    for w in which:
        globals()[w]()
