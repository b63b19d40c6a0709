"""A/B of the two generations of the two-barrier Jacobi body (ET_TUNE_EIG_THREADS 2001 / 2002; 3001 / 3002 add phase cycle
counters): accuracy of both against torch.linalg.eigh in float64 on a set of Gram matrices (1e6 synthetic pedestrians,
graded random, rank-deficient, tiny), timings, and where the cycles of a step go."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories

dev = torch.device("cuda")
lib = et.load_library()
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_000_000, seed=0))
G_o, G_p, _, _ = ops.gram_init(obs, pred)
o2, p2 = (x.to(dev) for x in synthetic_trajectories(181, seed=3))
g_o, g_p = ops.gram(o2, p2, True, True, True)
gen = torch.Generator().manual_seed(7)


def graded(m, decades, rank=None):
    q, _ = torch.linalg.qr(torch.randn(m, m, dtype=torch.float64, generator=gen))
    lam = torch.logspace(0, -decades, m, dtype=torch.float64)
    if rank is not None:
        lam[rank:] = 0
    return ((q * lam) @ q.T).to(dev)


cases = [("synthetic 1e6 pred", G_p), ("synthetic 1e6 obs", G_o), ("synthetic 181 pred", g_p), ("synthetic 181 obs", g_o),
         ("graded 24, 8 decades", graded(24, 8)), ("graded 16, 12 decades", graded(16, 12)), ("rank 5 of 24", graded(24, 3, rank=5)),
         ("identity 16", torch.eye(16, dtype=torch.float64, device=dev)), ("tiny 24 (1e-30 scale)", graded(24, 4) * 1e-30),
         ("pairs of equal eigenvalues 24", None)]
q, _ = torch.linalg.qr(torch.randn(24, 24, dtype=torch.float64, generator=gen))
lam = torch.tensor([float(2 ** -(i // 2)) for i in range(24)], dtype=torch.float64)
cases[-1] = (cases[-1][0], ((q * lam) @ q.T).to(dev))


def quality(G, U, S, k):
    G = 0.5 * (G + G.T)
    w, v = torch.linalg.eigh(G)
    w, v = w.flip(0), v.flip(1)
    lam = S * S
    res = float((G @ U - U * lam).abs().max() / w[0].abs().clamp_min(1e-300))
    orth = float((U.T @ U - torch.eye(k, dtype=torch.float64, device=G.device)).abs().max())
    srel = float(((S - w[:k].clamp_min(0).sqrt()).abs() / w[0].sqrt().clamp_min(1e-300)).max())
    return res, orth, srel


out = []
for name, G in cases:
    m = G.size(0)
    row = {"case": name, "m": m}
    ref = None
    for gen_id in (1, 2, 3):
        lib.et_tune(3, 2000 + gen_id)
        info = torch.zeros(2, dtype=torch.int32, device=dev)
        U, S, U64, S64 = ops.eig_basis(G, m, want64=True, info=info)
        res, orth, srel = quality(G, U64, S64, m)
        U6 = ops.eig_basis(G, 6, want64=True)[2]
        row[f"gen{gen_id}"] = {"sweeps_rotations": info.tolist(), "residual": res, "orthogonality": orth, "S_err_rel_to_S1": srel}
        if ref is None:
            ref = (U6, S64)
        else:
            row[f"projector6_gen{gen_id}_vs_gen1"] = float((U6 @ U6.T - ref[0] @ ref[0].T).norm())
            row[f"S_gen{gen_id}_vs_gen1_rel_to_S1"] = float(((S64 - ref[1]).abs() / ref[1][0].clamp_min(1e-300)).max())
    print(json.dumps(row), flush=True)
    out.append(row)
lib.et_tune(3, 0)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d = et.ETDescriptor(et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)).to(dev)


def init_cold():
    flush.fill_(1)
    d.parameter_initialization(obs, pred)


for gen_id in (1, 2, 3):
    lib.et_tune(3, 2000 + gen_id)
    t = {"generation": gen_id, "eig24_us": timed(lambda: ops.eig_basis(G_p, 6)), "eig16_us": timed(lambda: ops.eig_basis(G_o, 6)),
         "eig_pair_us": timed(lambda: ops.eig_basis_pair(G_o, G_p, 6))}
    ts = []
    for _ in range(12):
        flush.fill_(1); flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.parameter_initialization(obs, pred); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t["parameter_initialization_us"] = round(ts[len(ts) // 2], 1)
    print(json.dumps(t), flush=True)
names = ["setup", "param", "barrier1", "update", "barrier2", "rotating_steps", "idle_steps", "idle_step_cycles", "ordering", "output"]
for gen_id in (1, 2, 3):
    for tag, G in (("24x24", G_p), ("16x16", G_o)):
        if gen_id == 3 and tag == "16x16":
            continue
        lib.et_tune(3, 3000 + gen_id)
        info = torch.zeros(14, dtype=torch.int32, device=dev)
        ops.eig_basis(G, 6, info=info)
        v = info.tolist()
        print(json.dumps({"profile": f"generation {gen_id} {tag}", "sweeps_rotations": v[:2], **dict(zip(names, v[2:]))}), flush=True)
lib.et_tune(3, 0)
