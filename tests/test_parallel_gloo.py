"""CPU, world_size = 2, gloo: the host-side logic of the row-sharded paths (partitioning, packing, the single
all-reduce, global arg-min tie-breaking).  The numerical work is injected from the oracle, so no GPU is needed;
on the GPU box the same functions run with the CUDA backend (tests/test_gpu_multi.py, bench)."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleBackend:
    """Stands in for libet_b200.so on the CPU (test double built on oracle/et_oracle.py)."""

    def __init__(self):
        from oracle import et_oracle as O
        self.O = O

    def gram(self, obs, pred, ori, rot, sca):
        st = self.O.norm_params(obs.double(), ori, rot, sca)
        on = self.O.normalize(obs.double(), *st).reshape(obs.size(0), -1)
        pn = self.O.normalize(pred.double(), *st).reshape(pred.size(0), -1)
        return on.T @ on, pn.T @ pn

    def eig(self, G, k):
        lam, V = torch.linalg.eigh(G)
        lam, V = lam.flip(0)[:k], V.flip(1)[:, :k]
        idx = V.abs().argmax(dim=0)
        V = V * torch.sign(V[idx, torch.arange(k)])
        return V.float(), lam.clamp_min(0).sqrt().float()

    def new_workspace(self, l, d, k, device):
        ns, nc = l * d * k, l * k
        ws = types.SimpleNamespace()
        ws.flat = torch.zeros(ns + nc + l, dtype=torch.float64)
        ws.sums, ws.counts, ws.simsum = ws.flat[:ns].view(l, d, k), ws.flat[ns:ns + nc].view(l, k), ws.flat[ns + nc:]
        ws.simsum_last = torch.zeros(l, dtype=torch.float64)
        ws.status = torch.zeros(2, dtype=torch.int32)

        def reset():
            ws.flat.zero_(); ws.simsum_last.zero_(); ws.status.zero_()
        ws.reset = reset
        return ws

    def assign_accumulate(self, data, cent, acc):
        if int(acc.status[0]):
            return
        ms, lb = self.O.kmeans_assign(data, cent)
        k = cent.size(-1)
        onehot = torch.nn.functional.one_hot(lb, k).double()            # (l, n, k)
        acc.sums += data.double() @ onehot
        acc.counts += onehot.sum(dim=1)
        acc.simsum += ms.double().sum(dim=-1)

    def finalize(self, acc, old, new, tol):
        if int(acc.status[0]):
            return
        new.copy_((acc.sums / acc.counts[:, None, :]).float())
        err = float(((old.double() - new.double()) ** 2).sum())
        acc.simsum_last.copy_(acc.simsum)
        acc.flat.zero_()
        acc.status[1] += 1
        if err <= tol:
            acc.status[0] = 1

    def labels(self, data, cent):
        return self.O.kmeans_assign(data, cent)[1]

    def seed_step(self, data, cent, ncols):
        best, _ = self.O.kmeans_sim(data, cent[..., :ncols].contiguous()).max(dim=-1)
        val, idx = best.min(dim=-1)
        return val, idx


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")       # loopback: the box's hostname need not resolve
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out[("ready", rank)] = True
    try:
        from eigentrajectory_b200 import parallel as P
        from oracle import et_oracle as O
        torch.set_num_threads(1)
        be = OracleBackend()
        n = 2001                                              # odd: ragged shards
        obs, pred = O.synthetic_trajectories(n, seed=3)
        a, b = P.shard_bounds(n, rank, world)
        res = {}
        # ---- basis: one all-reduce, identical result on every rank, equal to the unsharded one ----
        Uo, So, Up, Sp = P.sharded_basis(obs[a:b], pred[a:b], 6, backend=be)
        full = O.parameter_initialization(obs.double(), pred.double(), 6)
        res["S_obs_err"] = float((So.double() - full["S_obs"]).abs().max() / full["S_obs"].max())
        res["P_pred_err"] = float((Up.double() @ Up.double().T - full["U_pred"] @ full["U_pred"].T).norm())
        res["U_pred"] = Up
        # ---- k-means over shards == k-means over everything ----
        g = torch.Generator().manual_seed(1234)
        data = (torch.randn(2, 6, n, generator=g) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous()
        first = 777
        cent0 = P.sharded_farthest_init(data[:, :, a:b].contiguous(), 20, first, a, backend=be)
        ref0 = O.kmeans_farthest_init(data, 20, first)
        res["init_equal"] = bool(torch.equal(cent0, ref0))
        labels, cent, n_iter, inertia = P.sharded_kmeans_fit(data[:, :, a:b].contiguous(), 20, n, cent0, max_iter=30,
                                                             sync_every=4, backend=be)
        ref_labels, ref_cent, ref_iter, ref_inertia = O.kmeans_fit(data, 20, centroids=ref0, max_iter=30)
        res["label_mismatch"] = int((labels != ref_labels[:, a:b]).sum())
        res["cent_err"] = float((cent - ref_cent).abs().max())
        res["iters"] = (n_iter, ref_iter)
        res["inertia"] = (inertia, float(ref_inertia))
        res["cent"] = cent
        res["mean"] = P.sharded_mean(torch.arange(a, b, dtype=torch.float32))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    from eigentrajectory_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_basis_and_kmeans_world2():
    from conftest import spawn_ranks
    mgr = mp.Manager()
    out = mgr.dict()
    spawn_ranks(_worker, 2, (out,), out=out)
    r0, r1 = out[0], out[1]
    for r in (r0, r1):
        assert r["S_obs_err"] < 1e-6 and r["P_pred_err"] < 1e-5
        assert r["init_equal"]
        assert r["label_mismatch"] <= 2 and r["cent_err"] < 1e-3      # fp64 sums vs the reference's fp32 sums
        assert r["iters"][0] == r["iters"][1]
        assert abs(r["inertia"][0] - r["inertia"][1]) < 1e-3 * abs(r["inertia"][1])
        assert abs(r["mean"] - 1000.0) < 1e-9
    assert torch.equal(r0["U_pred"], r1["U_pred"]) and torch.equal(r0["cent"], r1["cent"])
