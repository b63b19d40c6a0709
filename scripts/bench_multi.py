#!/usr/bin/env python
"""Config 5: row-sharded init path at N GPUs -- Gram pass + ONE all-reduce + eigen-solve, and k-means iterations with
one all-reduce each.  Launch with torchrun (one rank per GPU); prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_multi.py [--n-per-gpu 1250000]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigentrajectory_b200 as et                                   # noqa: E402
from eigentrajectory_b200 import ops, parallel as P                 # noqa: E402
from eigentrajectory_b200.synthetic import synthetic_trajectories   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n_per_gpu
    obs, pred = (x.to(dev) for x in synthetic_trajectories(n, seed=1000 + rank))

    def timed(fn, reps):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res = {"n_gpus": world, "n_per_gpu": n}
    res["basis_ms"] = timed(lambda: P.sharded_basis(obs, pred, 6), args.reps)
    res["gram_only_ms"] = timed(lambda: ops.gram(obs, pred, True, True, True), args.reps)
    Uo, So, Up, Sp = P.sharded_basis(obs, pred, 6)
    C = ops.project(obs, pred, Uo, Up)[1].unsqueeze(0).contiguous()
    a = rank * n
    cent0 = P.sharded_farthest_init(C, 20, 12345, a)
    res["seed_ms"] = timed(lambda: P.sharded_farthest_init(C, 20, 12345, a), 3)
    acc = ops.KMeansWorkspace(1, 6, 20, dev)
    nxt = torch.empty_like(cent0)

    def km_iter():
        ops.kmeans_assign(C, cent0, want_labels=False, want_maxsims=False, acc=acc, simsum=acc.simsum)
        if world > 1:
            dist.all_reduce(acc.flat)
        ops.kmeans_finalize(acc, cent0, nxt)
    res["kmeans_iter_ms"] = timed(km_iter, args.reps)
    if world > 1 and P.peer_exchange_available(dev):
        # whole sharded fit in one persistent kernel per rank, all-reduce over peer memory inside the kernel
        iters = 200
        res["kmeans_fused_iter_ms"] = timed(lambda: P.sharded_kmeans_fit_fused(C, 20, world * n, cent0, max_iter=iters, tol=-1.0),
                                            5) / iters
        res["points_per_s_kmeans_fused_iter"] = world * n / (res["kmeans_fused_iter_ms"] * 1e-3)
    out = (torch.empty_like(obs), torch.empty_like(pred), torch.empty((6, n), device=dev), torch.empty((6, n), device=dev))
    res["project_reconstruct_ms"] = timed(lambda: ops.project_reconstruct(obs, pred, Uo, Up, out=out), args.reps)
    res["traj_per_s_basis"] = world * n / (res["basis_ms"] * 1e-3)
    res["points_per_s_kmeans_iter"] = world * n / (res["kmeans_iter_ms"] * 1e-3)
    res["traj_per_s_project_reconstruct"] = world * n / (res["project_reconstruct_ms"] * 1e-3)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
