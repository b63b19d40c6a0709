#!/usr/bin/env python
"""Per-kernel roofline table for every op on the hot path (device-resident inputs, CUDA events).

    python scripts/bench_kernels.py [--out gpurun_out/kernels.json] [--cpu]

For each op: algorithmic bytes per unit (SURVEY.md section 8d / DESIGN.md), average launch duration
over rotating buffer sets larger than L2 (or an explicit L2 flush), achieved GB/s and the fraction of
the measured HBM peak.  --cpu also times the oracle (the reference's torch CPU path) on the host cores.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import eigentrajectory_b200 as et                      # noqa: E402
from eigentrajectory_b200 import ops                   # noqa: E402
from eigentrajectory_b200.synthetic import synthetic_trajectories   # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


_flush = None
_clean = None
FLUSH_MODE = "write+read"


def flush_l2():
    """Evict the op's tensors from the 126 MB L2: write a 256 MB buffer, then (mode "write+read") read another 256 MB
    so that the dirty lines of the write pass are written back BEFORE the timed region instead of inside it (a pure
    write flush leaves ~126 MB of dirty lines whose write-back the timed kernel would pay for: +15-20 us)."""
    global _flush, _clean
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        _clean = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    _flush.zero_()
    if FLUSH_MODE == "write+read":
        _clean.sum()


def time_op(fn, reps=20, warmup=3, flush=True):
    """Average / min device time of fn() in ms, CUDA events on the current stream; L2 flushed between launches."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.mean(times)), float(np.min(times))


def cpu_time(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--flush", default="write+read", choices=["write+read", "write"],
                    help="L2 flush between timed launches (see flush_l2)")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--n20", type=int, default=200_000, help="pedestrians for the S=20 ops")
    args = ap.parse_args()
    global FLUSH_MODE
    FLUSH_MODE = args.flush
    dev = torch.device("cuda")
    peak = peak_gbs()
    rows = []
    hp = et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)
    n, n20 = args.n, args.n20
    obs_h, pred_h = synthetic_trajectories(n, seed=0)
    obs, pred = obs_h.to(dev), pred_h.to(dev)
    desc = et.ETDescriptor(hp).to(dev)
    desc.parameter_initialization(obs, pred)
    Uo, Up = desc.U_obs_trunc.detach(), desc.U_pred_trunc.detach()

    def add(name, units, bytes_per_unit, avg_ms, min_ms, unit_name, note="", launches=1):
        gbs = units * bytes_per_unit / (avg_ms * 1e-3) / 1e9
        rows.append({"op": name, "units": units, "unit": unit_name, "algorithmic_bytes_per_unit": bytes_per_unit,
                     "avg_ms": avg_ms, "min_ms": min_ms, "units_per_s": units / (avg_ms * 1e-3), "achieved_gbs": gbs,
                     "frac_of_measured_hbm_peak": gbs / peak, "launches": launches, "note": note})
        print(json.dumps(rows[-1]), flush=True)

    # ---- headline op, all variants ----
    out = (torch.empty_like(obs), torch.empty_like(pred), torch.empty((6, n), device=dev), torch.empty((6, n), device=dev))
    for v in (1, 2, 3, 4):
        a, m = time_op(lambda: ops.project_reconstruct(obs, pred, Uo, Up, variant=v, out=out))
        add(f"project_reconstruct(variant={v})", n, 368, a, m, "trajectory")
    out_nc = (out[0], out[1], None, None)
    a, m = time_op(lambda: ops.project_reconstruct(obs, pred, Uo, Up, variant=2, out=out_nc))
    add("project_reconstruct(variant=2, coefficients not materialised)", n, 320, a, m, "trajectory")

    # ---- projection ----
    a, m = time_op(lambda: ops.project(obs, pred, Uo, Up))
    add("project (ETDescriptor.projection)", n, 236, a, m, "trajectory", "includes torch.empty of 5 outputs")

    # ---- reconstruction S=20, forward and backward ----
    o20, p20 = obs[:n20].contiguous(), pred[:n20].contiguous()
    _, _, state = ops.project(o20, p20, Uo, Up)
    C20 = (torch.randn(6, n20, 20, device=dev) * 0.5).contiguous()
    anchor = torch.randn(6, 20, device=dev)
    a, m = time_op(lambda: ops.reconstruct(C20, Up, state, anchor=anchor))
    add("reconstruct S=20 (+anchor)", n20, 2428, a, m, "pedestrian", "includes torch.empty of the output")
    rec = ops.reconstruct(C20, Up, state, anchor=anchor)
    gC = torch.empty_like(C20)
    lib = et.load_library()
    from eigentrajectory_b200._lib import ptr, stream_of, check

    def bwd():
        check(lib.et_reconstruct_bwd(ptr(rec), n20, 20, 6, 12, ptr(Up), 7, ptr(state[1]), ptr(state[2]), ptr(gC),
                                     stream_of(dev)), "bwd")
    a, m = time_op(bwd)
    add("reconstruct_bwd S=20", n20, 1920 + 480 + 20, a, m, "pedestrian")

    # ---- ADE / FDE ----
    a, m = time_op(lambda: ops.ade_fde(rec, p20))
    add("ade_fde S=20", n20, 2024, a, m, "pedestrian")

    # ---- eigen-basis ----
    Go = torch.zeros(16, 16, dtype=torch.float64, device=dev)
    Gp = torch.zeros(24, 24, dtype=torch.float64, device=dev)
    a, m = time_op(lambda: ops.gram(obs, pred, True, True, True, G_obs=Go, G_pred=Gp))
    add("gram (normalise + G_obs + G_pred, DMMA fp64)", n, 160, a, m, "trajectory")
    Gp1 = ops.gram(obs, pred, True, True, True)[1]
    a, m = time_op(lambda: ops.eig_basis(Gp1, 6), flush=False)
    add("eig_jacobi 24x24", 1, 24 * 24 * 8, a, m, "matrix", "latency-bound single block")
    small = ops.normalize(obs[:181].contiguous(), *ops.norm_params(obs[:181].contiguous(), True, True, False))
    a, m = time_op(lambda: ops.svd_small(small, 6), flush=False)
    add("svd_small 181x16 (one-sided Jacobi, 1 block)", 1, 181 * 16 * 4, a, m, "matrix", "latency-bound single block")
    a, m = time_op(lambda: desc.parameter_initialization(obs, pred))
    add("ETDescriptor.parameter_initialization (end to end)", n, 160 + 96 + 96, a, m, "trajectory",
        "norm_params + normalise(pred) + fused Gram + 2 eigen-solves", launches=5)

    # ---- k-means ----
    gen = torch.Generator().manual_seed(1234)
    data_h = (torch.randn(1, 6, n, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous()
    data = data_h.to(dev)
    km = et.BatchKMeans(n_clusters=20)
    np.random.seed(0)
    cent = km.initialize_centroids(data)
    acc = ops.KMeansWorkspace(1, 6, 20, dev)
    nxt = torch.empty_like(cent)

    def km_iter():
        ops.kmeans_assign(data, cent, want_labels=False, want_maxsims=False, acc=acc, simsum=acc.simsum)
        ops.kmeans_finalize(acc, cent, nxt)
    a, m = time_op(km_iter)
    add("kmeans assign+update iteration (HBM-cold)", n, 24, a, m, "point", "2 launches", launches=2)
    a, m = time_op(km_iter, flush=False)
    add("kmeans assign+update iteration (L2-resident, 24 MB)", n, 24, a, m, "point", "2 launches; data stays in the 126 MB L2",
        launches=2)
    a, m = time_op(lambda: ops.kmeans_assign(data, cent))
    add("kmeans get_labels (labels int64 + maxsims written)", n, 24 + 12, a, m, "point", "includes torch.empty")

    # the whole Lloyd loop in one persistent launch: tol < 0 never converges, so exactly 100 iterations run
    a, m = time_op(lambda: ops.kmeans_lloyd(data, cent, acc, 100, -1.0, want_labels=False), reps=5, flush=False)
    add("kmeans whole-fit kernel: per Lloyd iteration (100 iterations, 1 launch, L2-resident)", n, 24, a / 100, m / 100, "point",
        "et_kmeans_lloyd, in-kernel grid barriers")

    def seed():
        np.random.seed(0)
        km.initialize_centroids(data)
    a, m = time_op(seed, reps=5, flush=False)
    add("kmeans farthest-point init (K=20, one persistent launch; cold L2)", n, 24 * 19, a, m, "point", launches=2)

    def fit():
        np.random.seed(0)
        km.fit(data)
    a, m = time_op(fit, reps=3, warmup=1, flush=False)
    add(f"BatchKMeans.fit ({km.n_iter_} iterations + init + labels)", n, 24 * (km.n_iter_ + 20), a, m, "point")

    result = {"peak_gbs": peak, "gpu": torch.cuda.get_device_name(0), "l2_flush": FLUSH_MODE, "rows": rows}

    if args.cpu:
        from oracle import et_oracle as O
        threads = os.cpu_count()
        torch.set_num_threads(threads)
        cpu = []
        Uo_c, Up_c = Uo.cpu(), Up.cpu()

        def c(name, units, fn, reps=3):
            t = cpu_time(fn, reps)
            cpu.append({"op": name, "units": units, "seconds": t, "units_per_s": units / t, "cores": threads})
            print(json.dumps(cpu[-1]), flush=True)
        c("project_reconstruct", n, lambda: O.project_reconstruct(obs_h, pred_h, Uo_c, Up_c))
        c("projection", n, lambda: O.descriptor_projection(obs_h, pred_h, Uo_c, Up_c))
        c("parameter_initialization (2 thin SVDs incl. Vt)", n, lambda: O.parameter_initialization(obs_h, pred_h, 6), reps=2)
        st_c = O.norm_params(obs_h[:n20])
        C_c = C20.cpu()
        c("reconstruction S=20", n20, lambda: O.descriptor_reconstruction(C_c, Up_c, st_c), reps=2)
        rec_c = rec.cpu()
        c("ade+fde S=20 (two passes, as the reference)", n20,
          lambda: ((rec_c - pred_h[:n20]).norm(p=2, dim=-1).mean(dim=2).min(dim=0)[0],
                   (rec_c - pred_h[:n20]).norm(p=2, dim=-1)[:, :, -1].min(dim=0)[0]), reps=2)
        cent_c = cent.cpu()

        def cpu_iter():
            ms, lb = O.kmeans_assign(data_h, cent_c)
            return O.kmeans_update(data_h, lb, 20)
        c("kmeans assign+update iteration", n, cpu_iter, reps=2)
        result["cpu"] = cpu
        result["cpu_threads"] = threads

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(result, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
