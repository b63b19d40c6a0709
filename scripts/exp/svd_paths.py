"""One-sided Jacobi on the data (svd_small) vs fp64 Gram + eigen-solve, by size: wall time per call (CUDA events incl.
the Python wrapper) and distance of S / projector to a float64 torch SVD."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
obs, pred = synthetic_trajectories(4096, seed=7)
st = ops.norm_params(obs.to(dev), True, True, False)
xs = {8: ops.normalize(obs.to(dev), *st), 12: ops.normalize(pred.to(dev), *st)}

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for t in (8, 12):
    for n in (8, 32, 64, 181, 512, 2000):
        x = xs[t][:n].contiguous()
        M = x.reshape(n, -1).double().cpu()
        Uref, Sref, _ = torch.linalg.svd(M.T, full_matrices=False)
        k = min(6, n)
        def jac(): return ops.svd_small(x, k)
        def gram(): G, _ = ops.gram(x); return ops.eig_basis(G, k)
        out = {}
        for name, fn in (("jacobi", jac), ("gram+eig", gram)):
            if name == "jacobi" and not ops.svd_small_fits(n, t): continue
            U, S = fn()
            U, S = (U[0], S[0]) if name == "jacobi" else (U, S)
            P = U.double().cpu() @ U.double().cpu().T
            Pr = Uref[:, :k] @ Uref[:, :k].T
            out[name] = (timed(fn), float(((S.double().cpu() - Sref[:k]) / Sref[0]).abs().max()), float((P - Pr).norm()))
        print(f"T={t} N={n}: " + "  ".join(f"{k_}: {v[0]:6.1f} us S err {v[1]:.1e} proj err {v[2]:.1e}" for k_, v in out.items()))
