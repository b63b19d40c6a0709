"""CPU oracle for the EigenTrajectory hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, op for op, the arithmetic of the reference's descriptor /
normaliser / k-means / metric path as stateless functions on CPU torch tensors.
It is the *checker* for the CUDA kernels in ``eigentrajectory_b200/csrc``; nothing
in the product package imports it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

Parity status: PINNED to outputs of the reference itself.  ``tests/golden/make_golden.py``
imports the reference from ``/root/reference`` (CPU, torch 2.11) and freezes its
outputs into ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors (bit-exact for everything except the SVD,
whose LAPACK call is shared anyway).  The reference has no tests / golden vectors
of its own, and its third-party arithmetic (LAPACK gesdd via
``torch.linalg.svd``, MKL sgemm) is not version-pinned upstream; the pin is the
fixtures generated in this container.

All functions are dtype-generic: pass float64 tensors to obtain the fp64 "truth"
used to decide who is closer (ours or the reference's fp32 LAPACK).

Citations are ``path:line`` relative to the reference checkout.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Normaliser  (EigenTrajectory/normalizer.py)
# --------------------------------------------------------------------------------------


def norm_params(obs, use_ori=True, use_rot=True, use_sca=True):
    """Per-pedestrian origin / rotation / scale from the observed track.

    Follows ``EigenTrajectory/normalizer.py:17-28``: origin is the last observed
    frame, heading is atan2 of (last - third-from-last), rotation matrix rows are
    (cos, -sin), (sin, cos), scale is (1/||d||) * 2.
    Returns (ori (N,1,2) | None, rot (N,2,2) | None, sca (N,1,1) | None).
    """
    ori = rot = sca = None
    if use_ori:
        ori = obs[:, [-1]]
    if use_rot:
        d = obs[:, -1] - obs[:, -3]
        th = torch.atan2(d[:, 1], d[:, 0])
        c, s = th.cos(), th.sin()
        rot = torch.stack([torch.stack([c, -s], dim=1), torch.stack([s, c], dim=1)], dim=1)
    if use_sca:
        sca = 1.0 / (obs[:, -1] - obs[:, -3]).norm(p=2, dim=-1)[:, None, None] * 2
    return ori, rot, sca


def normalize(traj, ori, rot, sca):
    """``normalizer.py:42-51``: ((x - ori) @ R) * sca, each stage optional."""
    if ori is not None:
        traj = traj - ori
    if rot is not None:
        traj = traj @ rot
    if sca is not None:
        traj = traj * sca
    return traj


def denormalize(traj, ori, rot, sca):
    """``normalizer.py:53-62``: ((x / sca) @ R^T) + ori, each stage optional."""
    if sca is not None:
        traj = traj / sca
    if rot is not None:
        traj = traj @ rot.transpose(-1, -2)
    if ori is not None:
        traj = traj + ori
    return traj


def static_mask(obs, static_dist):
    """``EigenTrajectory/model.py:46,73``: moving iff ||(last - third_last)/2|| > static_dist."""
    return (obs[:, -1] - obs[:, -3]).div(2).norm(p=2, dim=-1) > static_dist


# --------------------------------------------------------------------------------------
# Descriptor  (EigenTrajectory/descriptor.py)
# --------------------------------------------------------------------------------------


def as_matrix(traj):
    """(N,T,2) -> the wide strided view M (2T, N) of ``descriptor.py:109``."""
    n = traj.shape[0]
    return traj.reshape(n, -1).T


def svd_basis(traj_norm, k):
    """Thin SVD of the wide view, truncated: ``descriptor.py:91-114``.

    Returns U (2T,k), S (k,), V (N,k) exactly as ``truncated_SVD`` does
    (LAPACK gesdd through torch.linalg.svd).
    """
    M = as_matrix(traj_norm)
    U, S, Vt = torch.linalg.svd(M, full_matrices=False)
    return U[:, :k], S[:k], Vt[:k, :].T


def project(traj_norm, U):
    """``descriptor.py:59-73``: C (k,N) = U^T M."""
    return U.T @ as_matrix(traj_norm)


def unproject(C, U, dim=2):
    """``descriptor.py:75-89``: (k,N) -> normalised trajectory (N,T,dim)."""
    t = U.shape[0] // dim
    return (U @ C).T.reshape(-1, t, dim)


def descriptor_projection(obs, pred, U_obs, U_pred, use_ori=True, use_rot=True, use_sca=True):
    """``ETDescriptor.projection`` (``descriptor.py:144-160``) as a pure function.

    Returns C_obs (k,N), C_pred (k,N)|None and the normaliser state (ori, rot, sca).
    """
    state = norm_params(obs, use_ori, use_rot, use_sca)
    C_obs = project(normalize(obs, *state), U_obs)
    C_pred = project(normalize(pred, *state), U_pred) if pred is not None else None
    return C_obs, C_pred, state


def descriptor_reconstruction(C_pred, U_pred, state, dim=2):
    """``ETDescriptor.reconstruction`` (``descriptor.py:162-176``).

    C_pred (k,N,S) -> (S,N,T,dim); one U @ C[:,:,s] product and one denormalise per
    sample, stacked on a new leading axis.
    """
    out = []
    for s in range(C_pred.shape[2]):
        out.append(denormalize(unproject(C_pred[:, :, s], U_pred, dim), *state))
    return torch.stack(out, dim=0)


def anchor_add(C_anchor, C_pred):
    """``ETAnchor.forward`` (``anchor.py:76-88``): (k,1,S) + (k,N,S)."""
    return C_anchor.unsqueeze(1) + C_pred


def parameter_initialization(obs, pred, k, use_ori=True, use_rot=True, use_sca=True):
    """``descriptor.py:116-142``: normalise, two truncated SVDs.

    Returns dict(U_obs, S_obs, U_pred, S_pred, obs_norm, pred_norm, state).
    """
    state = norm_params(obs, use_ori, use_rot, use_sca)
    obs_n, pred_n = normalize(obs, *state), normalize(pred, *state)
    U_o, S_o, _ = svd_basis(obs_n, k)
    U_p, S_p, _ = svd_basis(pred_n, k)
    return dict(U_obs=U_o, S_obs=S_o, U_pred=U_p, S_pred=S_p, obs_norm=obs_n, pred_norm=pred_n,
                state=state)


def project_reconstruct(obs, pred, U_obs, U_pred, use_ori=True, use_rot=True, use_sca=True):
    """The headline op: rank-k round trip of obs and pred with S=1.

    Shape of ``script/descriptor_evaluation.py:94-107`` (project, reconstruct,
    reshape, denormalise) on top of ``descriptor.py:144-160``.
    Returns (rec_obs (N,To,2), rec_pred (N,Tp,2), C_obs (k,N), C_pred (k,N)).
    """
    C_obs, C_pred, state = descriptor_projection(obs, pred, U_obs, U_pred, use_ori, use_rot, use_sca)
    rec_obs = denormalize(unproject(C_obs, U_obs), *state)
    rec_pred = denormalize(unproject(C_pred, U_pred), *state)
    return rec_obs, rec_pred, C_obs, C_pred


def rank_k_errors(obs, pred, kmax=12):
    """``script/descriptor_evaluation.py:87-112``: mean L2 error of the rank-k round trip.

    Normaliser is ori+rot without scale (``descriptor_evaluation.py:32``).
    Returns list of (k, obs_err, pred_err) plus the full U/S of both matrices.
    """
    state = norm_params(obs, True, True, False)
    A, B = as_matrix(normalize(obs, *state)), as_matrix(normalize(pred, *state))
    Uo, So, _ = torch.linalg.svd(A, full_matrices=False)
    Up, Sp, _ = torch.linalg.svd(B, full_matrices=False)
    rows = []
    n = obs.shape[0]
    for k in range(1, kmax + 1):
        Ao = Uo[:, :k] @ (Uo[:, :k].T @ A)
        Bo = Up[:, :k] @ (Up[:, :k].T @ B)
        ro = denormalize(Ao.T.reshape(n, -1, 2), *state)
        rp = denormalize(Bo.T.reshape(n, -1, 2), *state)
        rows.append((k, (ro - obs).norm(p=2, dim=-1).mean().item(),
                     (rp - pred).norm(p=2, dim=-1).mean().item()))
    return rows, (Uo, So, Up, Sp)


# --------------------------------------------------------------------------------------
# k-means  (EigenTrajectory/kmeans.py)
# --------------------------------------------------------------------------------------


def kmeans_sim(a, b):
    """``kmeans.py:59-76``: negative squared distance y = 2 a^T b - |a|^2 - |b|^2.

    a (..., d, m), b (..., d, n) -> (..., m, n); same in-place op order.
    """
    y = a.transpose(-2, -1) @ b
    y.mul_(2)
    y.sub_(a.pow(2).sum(dim=-2)[..., :, None])
    y.sub_(b.pow(2).sum(dim=-2)[..., None, :])
    return y


def kmeans_assign(data, centroids):
    """``kmeans.py:143-158``: (maxsim, label) = max over clusters of kmeans_sim."""
    return kmeans_sim(data, centroids).max(dim=-1)


def kmeans_update(data, labels, n_clusters):
    """``kmeans.py:160-184``: masked sum / count; empty cluster -> NaN (0/0)."""
    mask = torch.stack([labels == i for i in range(n_clusters)], dim=-1)
    return (data.unsqueeze(-1) * mask.unsqueeze(-3)).sum(dim=-2) / mask.sum(dim=-2, keepdim=True)


def kmeans_farthest_init(data, n_clusters, first_index):
    """``kmeans.py:78-112``: deterministic farthest-point seeding.

    The first centroid is ``data[..., first_index]`` (the reference draws it from
    ``np.random.randint(n_data)``, ``kmeans.py:93``); each next one is the point whose
    best similarity to the chosen set is lowest.  data (l, d, N) only.
    """
    assert data.dim() == 3
    l, d, _ = data.shape
    cent = torch.zeros(l, d, n_clusters, dtype=data.dtype)
    cent[..., 0] = data[..., first_index]
    rows = torch.arange(l)
    for i in range(1, n_clusters):
        best, _ = kmeans_sim(data, cent[..., :i].contiguous()).max(dim=-1)
        idx = best.argmin(dim=-1)
        cent[..., i] = data[rows, :, idx]
    return cent


def kmeans_fit(data, n_clusters, centroids=None, first_index=None, max_iter=100, tol=1e-4,
               trace=None):
    """``kmeans.py:200-259`` with n_redo=1: Lloyd iterations until sum((c-c')^2) <= tol.

    Returns (labels (l,N) int64, centroids (l,d,K), n_iter, inertia).  ``trace`` (a list)
    receives (centroids_in, labels, centroids_out) per iteration for lock-step tests.
    """
    if centroids is None:
        if first_index is None:
            first_index = np.random.randint(data.shape[-1])
        centroids = kmeans_farthest_init(data, n_clusters, first_index).clone()
    labels = inertia = None
    it = 0
    for it in range(1, max_iter + 1):
        maxsims, labels = kmeans_assign(data, centroids)
        new_c = kmeans_update(data, labels, n_clusters)
        err = (centroids - new_c).pow(2).sum()
        if trace is not None:
            trace.append((centroids.clone(), labels.clone(), new_c.clone()))
        centroids = new_c
        inertia = (-maxsims).mean()
        if err <= tol:
            break
    return labels, centroids, it, inertia


# --------------------------------------------------------------------------------------
# Metrics  (utils/metrics.py)
# --------------------------------------------------------------------------------------


def ade_fde(pred, gt):
    """``utils/metrics.py:73-102``: min-over-samples ADE and FDE per pedestrian.

    pred (S,N,T,2), gt (N,T,2) or (1,N,T,2) -> (ade (N,), fde (N,), argmin_fde (N,)).
    """
    dist = (pred - gt).norm(p=2, dim=-1)            # (S,N,T)
    ade = dist.mean(dim=2).min(dim=0)[0]
    fde, arg = dist[:, :, -1].min(dim=0)
    return ade, fde, arg


def tcc(pred, gt):
    """``utils/metrics.py:105-130``: temporal correlation coefficient of the best-FDE sample with the ground truth.

    pred (S,N,T,2), gt (N,T,2) or (1,N,T,2) -> (N,)."""
    gt = gt.squeeze(dim=0) if gt.dim() == 4 else gt
    temp = (pred - gt).norm(p=2, dim=-1)
    pred_best = pred[temp[:, :, -1].argmin(dim=0), range(pred.size(1)), :, :]
    stack = torch.stack([pred_best, gt], dim=0).permute(3, 1, 0, 2)          # (xy, N, {pred, gt}, T)
    cov = stack - stack.mean(dim=-1, keepdim=True)
    factor = 1 / (cov.shape[-1] - 1)
    cov = factor * cov @ cov.transpose(-1, -2)
    std = cov.diagonal(offset=0, dim1=-2, dim2=-1).sqrt()
    corr = (cov / std.unsqueeze(-1) / std.unsqueeze(-2)).clamp(-1, 1)
    corr[torch.isnan(corr)] = 0
    return corr[:, :, 0, 1].mean(dim=0)


def col(pred, num_interp=4, thres=0.2):
    """``utils/metrics.py:133-155``: per-pedestrian collision rate (percent of samples in which the pedestrian comes
    within ``thres`` of another one of the same scene during the first 3.25 interpolated steps).  pred (S,N,T,2) -> (N,)."""
    pred = pred.permute(0, 2, 1, 3)
    fp = pred[:, [0], :, :]
    rel = pred[:, 1:] - pred[:, :-1]
    rel_dense = rel.div(num_interp).unsqueeze(dim=2).repeat_interleave(repeats=num_interp, dim=2).contiguous()
    rel_dense = rel_dense.reshape(pred.size(0), num_interp * (pred.size(1) - 1), pred.size(2), pred.size(3))
    dense = torch.cat([fp, rel_dense], dim=1).cumsum(dim=1)
    m = dense[:, :3 * num_interp + 2].unsqueeze(dim=2).repeat_interleave(repeats=pred.size(2), dim=2)
    m = (m - m.transpose(2, 3)).norm(p=2, dim=-1)
    m = m.add(torch.eye(n=pred.size(2))[None, None, :, :].to(m.dtype)).min(dim=1)[0].lt(thres)
    return m.sum(dim=1).gt(0).type(pred.dtype).mean(dim=0).mul(100)


def forward_losses(C_pred, C_pred_gt, recon, pred_gt):
    """``EigenTrajectory/model.py:119-123``: the three training-loss scalars."""
    e_c = (C_pred - C_pred_gt.unsqueeze(-1)).norm(p=2, dim=0)
    e_d = (recon - pred_gt.unsqueeze(0)).norm(p=2, dim=-1)
    return (e_c.min(dim=-1)[0].mean(), e_d.mean(dim=-1).min(dim=0)[0].mean(),
            e_d[:, :, -1].min(dim=0)[0].mean())


# --------------------------------------------------------------------------------------
# Synthetic workload -- shared by tests and bench's CPU leg
# --------------------------------------------------------------------------------------


def synthetic_trajectories(n, seed=0, t_obs=8, t_pred=12, dtype=torch.float32):
    """Seeded pedestrian tracks: constant-turn-rate walkers with velocity noise.

    p0 ~ U(-10,10)^2, heading ~ U(0,2pi), speed ~ U(0.2,0.8) m/frame (never static),
    turn rate ~ N(0,0.05^2), per-frame velocity noise N(0,0.03^2).
    Returns obs (n,t_obs,2), pred (n,t_pred,2), contiguous.
    """
    g = torch.Generator().manual_seed(seed)
    T = t_obs + t_pred
    p0 = (torch.rand(n, 1, 2, generator=g) * 20 - 10)
    th0 = torch.rand(n, 1, generator=g) * (2 * np.pi)
    v = torch.rand(n, 1, generator=g) * 0.6 + 0.2
    om = torch.randn(n, 1, generator=g) * 0.05
    t = torch.arange(T, dtype=torch.float32)[None, :]
    ang = th0 + om * t
    vel = torch.stack([v * ang.cos(), v * ang.sin()], dim=-1) + torch.randn(n, T, 2, generator=g) * 0.03
    traj = (p0 + vel.cumsum(dim=1)).to(dtype)
    return traj[:, :t_obs].contiguous(), traj[:, t_obs:].contiguous()


# --------------------------------------------------------------------------------------
# Dataset preprocessing  (utils/dataloader.py)
# --------------------------------------------------------------------------------------


def dataset_parse(text, delim="\t"):
    """``read_file`` (``utils/dataloader.py:121-132``): every line stripped, split on ``delim``, fields through
    ``float()``.  Returns an (n_rows, 4) float64 array."""
    return np.asarray([[float(tok) for tok in line.strip().split(delim)] for line in text.splitlines()], dtype=np.float64)


def dataset_windows(rows, obs_len=8, pred_len=12, skip=1, threshold=0.02, min_ped=1):
    """The per-file body of ``TrajectoryDataset.__init__`` (``utils/dataloader.py:189-226``) and ``poly_fit``
    (``:135-151``), restated with numpy group operations instead of per-pedestrian masks.

    rows (n, 4) float64 ``frame, ped, x, y`` -> (traj (N, T, 2) float32, non_linear (N,) float32,
    num_peds_in_seq (n_seq,) int64).  A window = ``T = obs_len + pred_len`` consecutive distinct frames starting at
    frame index 0, skip, 2*skip, ...; a pedestrian is kept when its first / last rows in the window sit on the
    window's first / last frame (it must then have exactly T rows); coordinates are rounded with ``np.around(.., 4)``;
    a window is kept when MORE than ``min_ped`` pedestrians are."""
    T = obs_len + pred_len
    frames = np.unique(rows[:, 0])
    fidx = np.searchsorted(frames, rows[:, 0])
    order = np.argsort(fidx, kind="stable")                      # frame-major, file order inside a frame
    rows, fidx = rows[order], fidx[order]
    n_seq = int(np.ceil((len(frames) - T + 1) / skip))
    t = np.linspace(0, pred_len - 1, pred_len)
    trajs, flags, counts = [], [], []
    for start in range(0, n_seq * skip + 1, skip):
        sel = (fidx >= start) & (fidx < start + T)
        if not sel.any():
            raise ValueError("need at least one array to concatenate")
        win, wf = rows[sel], fidx[sel]
        peds, inv = np.unique(win[:, 1], return_inverse=True)
        kept, kept_flags = [], []
        for p in range(len(peds)):                               # ascending pedestrian id, as np.unique orders them
            mine = np.flatnonzero(inv == p)
            if wf[mine[-1]] - wf[mine[0]] + 1 != T:
                continue
            if len(mine) != T:
                raise ValueError("could not broadcast input array")      # what numpy raises in the reference
            xy = np.around(win[mine, 2:4], decimals=4).T         # (2, T)
            kept.append(xy)
            res = sum(np.polyfit(t, xy[axis, -pred_len:], 2, full=True)[1] for axis in (0, 1))
            kept_flags.append(1.0 if res >= threshold else 0.0)
        if len(kept) > min_ped:
            trajs.append(np.stack(kept))                         # (P, 2, T)
            flags += kept_flags
            counts.append(len(kept))
    if not trajs:
        return (np.zeros((0, T, 2), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.int64))
    traj = np.concatenate(trajs, axis=0).transpose(0, 2, 1).astype(np.float32)
    return traj, np.asarray(flags, dtype=np.float32), np.asarray(counts, dtype=np.int64)


# --------------------------------------------------------------------------------------
# k-means++ seeding by D^2 sampling  (what sklearn does inside anchor.py:65-71; third-party, restated from the
# published algorithm -- Arthur & Vassilvitskii 2007 with sklearn's greedy local trials -- NOT bit-pinned to sklearn)
# --------------------------------------------------------------------------------------


def kmeans_d2_seeding(data, n_clusters, uniform):
    """Specification of ``et_kmeans_d2_init`` in numpy: data (d, N) float32, uniform (K, trials) float64 in [0, 1).

    Centre 0 = column floor(u[0,0] * N).  Step i: total = sum of D^2 (float64); candidate j = first column whose
    inclusive cumulative D^2 exceeds u[i,j] * total; the candidate with the lowest potential sum(min(D^2, dist^2)) wins
    (lowest j on ties); D^2 = min(D^2, dist^2 to the winner).  dist^2 = ascending fp32 FMA chain of (x - c)^2.
    Returns (centroids (d, K) float32, chosen column indices (K,))."""
    x = np.asarray(data, dtype=np.float32)
    d, n = x.shape
    u = np.asarray(uniform, dtype=np.float64)
    trials = u.shape[1]

    def dist2(c):
        acc = np.zeros(n, dtype=np.float32)
        for r in range(d):
            df = (x[r] - np.float32(c[r])).astype(np.float32)
            acc = (df.astype(np.float64) * df.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)   # fused multiply-add
        return acc

    first = min(n - 1, int(u[0, 0] * n))
    picks = [first]
    d2 = dist2(x[:, first])
    for i in range(1, n_clusters):
        total = d2.astype(np.float64).sum()
        if total <= 0:
            picks.append(0)
            d2 = np.minimum(d2, dist2(x[:, 0]))
            continue
        cum = np.cumsum(d2.astype(np.float64))
        cands = [min(n - 1, int(np.searchsorted(cum, u[i, j] * total, side="right"))) for j in range(trials)]
        pots = [np.minimum(d2, dist2(x[:, c])).astype(np.float64).sum() for c in cands]
        best = int(np.argmin(pots))
        picks.append(cands[best])
        d2 = np.minimum(d2, dist2(x[:, cands[best]]))
    return x[:, picks].copy(), np.asarray(picks)


# --------------------------------------------------------------------------------------
# Jacobi eigen-solve of the Gram matrix: a numpy MODEL of the kernel's arithmetic (eig_jacobi_fast2 in
# eigentrajectory_b200/csrc/et_svd.cu), not a restatement of the reference -- the reference calls LAPACK
# (descriptor.py:110) and the parity statement for the bases is made against the golden fp32 / fp64 SVDs.  The model
# pins what the kernel's shortened rotation chain relies on: a rotation built from ~21-bit reciprocal-square-root seeds
# is orthogonal to rounding, and the solve needs the same sweeps as one with exactly computed angles.
# --------------------------------------------------------------------------------------


def _rsqrt_seed(x, extra_rel_error=0.0):
    """A stand-in for rsqrt.approx.ftz.f64: only the upper 32 bits of the operand are looked at and only the upper 32
    bits of the result are produced (relative error ~2^-21); ``extra_rel_error`` perturbs it further."""
    import struct
    hi = struct.unpack("<d", struct.pack("<Q", struct.unpack("<Q", struct.pack("<d", float(x)))[0] & 0xFFFFFFFF00000000))[0]
    r = (1.0 / np.sqrt(hi)) * (1.0 + extra_rel_error)
    return struct.unpack("<d", struct.pack("<Q", struct.unpack("<Q", struct.pack("<d", float(r)))[0] & 0xFFFFFFFF00000000))[0]


def jacobi_rotation_short_chain(app, aqq, apq, seed_error=0.0):
    """(c, s) of the inner Jacobi rotation of the pair (p, q) as eig_jacobi_fast2 computes it: 1/h from the seed and one
    Newton step, c^2 = 1/2 + |dd| / 2h, (c~, s~) = (c^2, +-o / 2h) * seed(c^2), normalised by 1 - d/2 + 3 d^2 / 8."""
    o, dd = 2.0 * apq, aqq - app
    x = dd * dd + o * o
    rh = _rsqrt_seed(x, seed_error)
    rh = rh * 0.5 * (-x * rh * rh + 1.0) + rh
    c2 = 0.5 * abs(dd) * rh + 0.5
    rc = _rsqrt_seed(c2, -seed_error)
    ct, st = c2 * rc, (0.5 if dd >= 0.0 else -0.5) * o * rh * rc
    d = ct * ct + (st * st - 1.0)
    n = 1.0 + d * (-0.5 + 0.375 * d)
    return ct * n, st * n


def jacobi_rotation_exact(app, aqq, apq, seed_error=0.0):
    """The textbook form (tangent of the inner angle, IEEE sqrt and division)."""
    o, dd = 2.0 * apq, aqq - app
    h = np.sqrt(dd * dd + o * o)
    t = o / (dd + (h if dd >= 0.0 else -h))
    c = 1.0 / np.sqrt(t * t + 1.0)
    return c, t * c


def jacobi_eig_model(G, k, rotation=jacobi_rotation_short_chain, rel=1e-12, max_sweeps=60, seed_error=0.0):
    """Parallel-ordered cyclic Jacobi as the kernel runs it: round-robin pairs (position 0 fixed), all rotations of a step
    from the matrix BEFORE the step, A <- J^T A J on ONE symmetric copy, the rotated pair stored as computed; rotate iff
    a_pq^2 > rel^2 |a_pp a_qq| and above the floor (1e-18 max diagonal)^2.  Returns (U (m, k), S (k), rotations per sweep,
    worst |c^2 + s^2 - 1|); columns ordered by descending eigenvalue, largest-magnitude component positive."""
    A = 0.5 * (np.asarray(G, dtype=np.float64) + np.asarray(G, dtype=np.float64).T)
    m = A.shape[0]
    assert m % 2 == 0
    V = np.eye(m)
    floor2 = (1e-18 * np.abs(np.diag(A)).max()) ** 2
    counts, worst = [], 0.0
    for _ in range(max_sweeps):
        nrot = 0
        for step in range(m - 1):
            def player(pos):
                return 0 if pos == 0 else 1 + (pos - 1 + step) % (m - 1)
            J = np.eye(m)
            for pi in range(m // 2):
                p, q = sorted((player(pi), player(m - 1 - pi)))
                app, aqq, apq = A[p, p], A[q, q], A[p, q]
                if apq * apq > floor2 and apq * apq > rel * rel * abs(app * aqq):
                    c, s = rotation(app, aqq, apq, seed_error)
                    worst = max(worst, abs(c * c + s * s - 1.0))
                    J[p, p], J[q, q], J[p, q], J[q, p] = c, c, s, -s
                    nrot += 1
            if nrot:
                A = J.T @ A @ J
                A = 0.5 * (A + A.T)
                V = V @ J
        counts.append(nrot)
        if nrot == 0:
            break
    lam = np.diag(A)
    order = np.argsort(-lam, kind="stable")[:k]
    U = V[:, order]
    U = U * np.where(U[np.abs(U).argmax(axis=0), np.arange(k)] < 0.0, -1.0, 1.0)
    return U, np.sqrt(np.maximum(lam[order], 0.0)), counts, worst
