"""Row-sharded (one process per GPU) versions of the two stages that need an exchange.

Trajectories are partitioned into contiguous row blocks, one per rank.  Projection, reconstruction
and ADE/FDE need no communication at all.  The eigen-basis needs ONE all-reduce (the fp64 Gram
accumulators, 832 doubles); k-means needs one all-reduce of ``l * (d*K + K + 1)`` doubles per Lloyd
iteration plus, for the reference's farthest-point seeding, one (value, index) exchange and one
broadcast per step.  All of them are a few KB: latency-bound, enqueued on the compute stream through
``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU unit tests).

The numerical work is injected through a small ``backend`` object so that the host-side logic
(partitioning, packing, collectives, tie-breaking) is unit-testable without a GPU; the default
backend is the CUDA library.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n, rank, world):
    """Contiguous, balanced row block [start, end) of rank ``rank`` out of ``world``."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class CudaBackend:
    """Default compute backend: libet_b200.so through ``ops``."""

    def gram(self, obs, pred, ori, rot, sca):
        return ops.gram(obs, pred, ori, rot, sca)

    def eig(self, G, k):
        return ops.eig_basis(G, k)

    def eig_pair(self, G_a, G_b, k):
        return ops.eig_basis_pair(G_a, G_b, k)

    def new_workspace(self, l, d, k, device):
        return ops.KMeansWorkspace(l, d, k, device)

    def assign_accumulate(self, data, cent, acc, row_offset=0, n_global=0):
        ops.kmeans_assign(data, cent, want_labels=False, want_maxsims=False, acc=acc, simsum=acc.simsum, status=acc.status,
                          row_offset=row_offset, n_global=n_global)

    def finalize(self, acc, old, new, tol):
        ops.kmeans_finalize(acc, old, new, tol=tol, use_status=True)

    def labels(self, data, cent, row_offset=0, n_global=0):
        return ops.kmeans_assign(data, cent, want_labels=True, want_maxsims=False, row_offset=row_offset, n_global=n_global)[1]

    def seed_candidate(self, data, cent, ncols, row_offset, out=None, n_global=0):
        return ops.kmeans_seed_candidate(data, cent, ncols, row_offset, out=out, n_global=n_global)

    def seed_fetch(self, data, row_offset, gkey, out=None):
        return ops.kmeans_seed_fetch(data, row_offset, gkey, out=out)

    def seed_step(self, data, cent, ncols):
        return ops.kmeans_seed_step(data, cent, ncols)


def _all_reduce(t, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_basis(obs_shard, pred_shard, k, ori=True, rot=True, sca=True, group=None, backend=None):
    """Eigen-bases of the row-sharded init set: local fused Gram pass, ONE all-reduce, replicated eigen-solve.

    Every rank returns the same (U_obs (2T_obs,k), S_obs, U_pred (2T_pred,k), S_pred): the Jacobi solve is
    deterministic, so no broadcast is needed."""
    backend = backend or CudaBackend()
    G_obs, G_pred = backend.gram(obs_shard, pred_shard, ori, rot, sca)
    no, np_ = G_obs.numel(), G_pred.numel()
    packed = torch.cat([G_obs.reshape(-1), G_pred.reshape(-1)])        # one buffer => one collective
    _all_reduce(packed, group)
    G_obs = packed[:no].reshape(G_obs.shape).contiguous()
    G_pred = packed[no:no + np_].reshape(G_pred.shape).contiguous()
    if hasattr(backend, "eig_pair"):
        (U_obs, S_obs), (U_pred, S_pred) = backend.eig_pair(G_obs, G_pred, k)     # both solves in one launch
    else:
        U_obs, S_obs = backend.eig(G_obs, k)
        U_pred, S_pred = backend.eig(G_pred, k)
    return U_obs, S_obs, U_pred, S_pred


def sharded_farthest_init(data_shard, n_clusters, first_global_index, row_offset, group=None, backend=None, n_total=None):
    """The reference's farthest-point seeding (kmeans.py:78-112) over row shards.

    ``data_shard`` (l,d,n_local) holds global columns [row_offset, row_offset + n_local).  Per step: local
    candidate (lowest best-similarity, lowest index on ties) -> all-gather of (value, global index) -> the
    global winner (lowest value, then lowest global index) -> its coordinates summed in from the owner.
    ``n_total`` (global column count): when given, the similarity arithmetic follows the global column numbering, i.e.
    the picks are the same bits as seeding the unsharded tensor."""
    backend = backend or CudaBackend()
    rank, world = _world(group)
    l, d, n_local = data_shard.shape
    dev = data_shard.device
    cent = torch.zeros((l, d, n_clusters), device=dev, dtype=data_shard.dtype)

    def fetch(global_idx):
        """(l,d) coordinates of global column ``global_idx`` (l,) from whichever rank owns each."""
        local = global_idx - row_offset
        mine = (local >= 0) & (local < n_local)
        safe = local.clamp(0, max(n_local - 1, 0))
        pts = data_shard[torch.arange(l, device=dev), :, safe] if n_local > 0 else torch.zeros((l, d), device=dev)
        pts = torch.where(mine[:, None], pts, torch.zeros_like(pts)).double()
        _all_reduce(pts, group)                     # exactly one rank contributes a non-zero row
        return pts.to(data_shard.dtype)

    first = torch.full((l,), int(first_global_index), dtype=torch.int64, device=dev)
    if hasattr(backend, "seed_candidate"):      # (every rank must take the same branch: the collectives differ)
        # device-only step: candidate key -> ONE int64 MIN all-reduce -> owner's coordinates -> ONE SUM all-reduce
        # (six small launches and two collectives per step, nothing decoded on the host)
        gkey = first ^ torch.iinfo(torch.int64).min           # key of "global column first", similarity bits zero
        coords = torch.empty((l, d), dtype=torch.float64, device=dev)
        for i in range(n_clusters):
            if i > 0:
                gkey = (backend.seed_candidate(data_shard, cent, i, row_offset, gkey, n_global=int(n_total)) if n_total
                        else backend.seed_candidate(data_shard, cent, i, row_offset, gkey))
                if world > 1:
                    dist.all_reduce(gkey, op=dist.ReduceOp.MIN, group=group)
            backend.seed_fetch(data_shard, row_offset, gkey, coords)
            _all_reduce(coords, group)
            cent[:, :, i].copy_(coords)
        return cent
    cent[:, :, 0] = fetch(first)
    big = torch.iinfo(torch.int64).max
    for i in range(1, n_clusters):
        if n_local > 0:
            val, idx = backend.seed_step(data_shard, cent, i)
            gidx = idx + row_offset
        else:
            val = torch.full((l,), float("inf"), device=dev)
            gidx = torch.full((l,), big, dtype=torch.int64, device=dev)
        if world > 1:
            vals = [torch.empty_like(val) for _ in range(world)]
            idxs = [torch.empty_like(gidx) for _ in range(world)]
            dist.all_gather(vals, val, group=group)
            dist.all_gather(idxs, gidx, group=group)
            vals, idxs = torch.stack(vals), torch.stack(idxs)            # (world, l)
            best = vals.min(dim=0, keepdim=True)[0]
            cand = torch.where(vals == best, idxs, torch.full_like(idxs, big))
            gidx = cand.min(dim=0)[0]
        cent[:, :, i] = fetch(gidx)
    return cent


class PeerExchange:
    """Exchange buffers of the in-kernel all-reduce of ``et_kmeans_lloyd_sharded``: one symmetric-memory allocation per
    rank, mapped into every other rank of the node (NVLink / NVSwitch peer access), plus the call counter that keeps
    the flag stamps monotonic.  Needs one process per GPU on one node and a NCCL process group."""

    _cache = {}

    def __init__(self, l, d, k, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import load
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(load().et_kmeans_exchange_bytes(l, d, k, self.world))
        if nbytes == 0:
            raise RuntimeError(f"peer exchange supports at most 16 ranks (world = {self.world})")
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.stamp = 1
        torch.cuda.synchronize(device)
        dist.barrier(group)                 # every buffer is zero before anybody stores into it

    @classmethod
    def get(cls, l, d, k, device, group=None):
        key = (l, d, k, str(device), id(group))
        if key not in cls._cache:
            cls._cache[key] = cls(l, d, k, device, group)
        return cls._cache[key]

    def next_stamp(self, max_iter):
        s = self.stamp
        self.stamp += max_iter + 2
        return s


def peer_exchange_available(device, group=None):
    """True when the fused (peer-memory) sharded fit can run: CUDA tensors, NCCL group, more than one rank."""
    if not (dist.is_available() and dist.is_initialized()) or torch.device(device).type != "cuda":
        return False
    if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
        return False
    return True


def sharded_farthest_init_fused(data_shard, n_clusters, first_global_index, row_offset, n_total, group=None):
    """``sharded_farthest_init`` with all K - 1 steps AND their per-step candidate exchange inside one persistent kernel
    per rank (``et_kmeans_farthest_init_sharded``, peer memory): no NCCL call and no host round trip per step.  The
    centroids are identical on every rank and equal to what the unsharded ``et_kmeans_farthest_init`` picks."""
    l, d, n_local = data_shard.shape
    dev = data_shard.device
    ex = PeerExchange.get(l, d, n_clusters, dev, group)
    return ops.kmeans_farthest_init_sharded(data_shard, n_clusters, first_global_index, row_offset, n_total, ex.rank, ex.world,
                                            ex.peers, ex.next_stamp(n_clusters))


def sharded_kmeans_fit_fused(data_shard, n_clusters, n_total, centroids, max_iter=100, tol=1e-4, group=None, row_offset=None):
    """``sharded_kmeans_fit`` with the whole Lloyd loop AND its per-iteration all-reduce inside one persistent kernel per
    rank (``et_kmeans_lloyd_sharded``): the ranks exchange their folded records through peer memory, no NCCL call and
    no host round trip per iteration.  Same return values; centroids and iteration count are identical on all ranks by
    construction (every rank adds the same records in rank order).  ``row_offset`` (global index of this shard's first
    column): when given, labels are the same bits as a fit of the unsharded (l, d, n_total) tensor would assign."""
    l, d, n_local = data_shard.shape
    dev = centroids.device
    ex = PeerExchange.get(l, d, n_clusters, dev, group)
    acc = ops.KMeansWorkspace(l, d, n_clusters, dev)
    labels, final = ops.kmeans_lloyd_sharded(data_shard, centroids.contiguous(), acc, max_iter, tol, ex.rank, ex.world,
                                             ex.peers, ex.next_stamp(max_iter), row_offset=row_offset or 0,
                                             n_global=n_total if row_offset is not None else 0)
    _, n_iter = (int(v) for v in acc.status.tolist())
    if labels is None:
        labels = torch.zeros((l, 0), dtype=torch.int64, device=dev)
    inertia = float(-(acc.simsum_last / n_total).mean())
    return labels, final, n_iter, inertia


def sharded_kmeans_fit(data_shard, n_clusters, n_total, centroids, max_iter=100, tol=1e-4, sync_every=8, group=None,
                       backend=None, row_offset=None):
    """Lloyd iterations (kmeans.py:200-259, n_redo = 1) over row shards: per iteration one fused
    assign+accumulate pass on local rows and ONE all-reduce of the packed (sums, counts, sum of similarities).

    Returns (labels of the local rows (l,n_local) int64, centroids (l,d,K) -- identical on all ranks --,
    iterations executed, inertia).  ``row_offset`` (global index of this shard's first column): when given, the
    assignment arithmetic follows the global column numbering, i.e. the labels are the same bits as a fit of the
    unsharded (l, d, n_total) tensor would assign."""
    backend = backend or CudaBackend()
    where = {} if row_offset is None else {"row_offset": int(row_offset), "n_global": int(n_total)}
    l, d, n_local = data_shard.shape
    dev = data_shard.device
    acc = backend.new_workspace(l, d, n_clusters, dev)
    bufs = [centroids.contiguous().clone(), torch.empty_like(centroids)]
    acc.reset()
    done = n_iter = 0
    while done < max_iter:
        chunk = min(sync_every, max_iter - done)
        for j in range(done, done + chunk):
            cur, nxt = bufs[j % 2], bufs[(j + 1) % 2]
            backend.assign_accumulate(data_shard, cur, acc, **where)
            _all_reduce(acc.flat, group)            # sums | counts | sum of similarities: one in-place collective
            backend.finalize(acc, cur, nxt, tol)
        done += chunk
        converged, n_iter = (int(v) for v in acc.status.tolist())
        if converged:
            break
    final, before = bufs[n_iter % 2], bufs[(n_iter - 1) % 2]
    labels = backend.labels(data_shard, before, **where) if n_local > 0 else torch.zeros((l, 0), dtype=torch.int64, device=dev)
    inertia = float(-(acc.simsum_last / n_total).mean())
    return labels, final.clone(), n_iter, inertia


def sharded_mean(values, group=None):
    """Mean of a per-pedestrian metric over all ranks: one 2-scalar all-reduce (sum, count)."""
    v = torch.as_tensor(values)
    pair = torch.tensor([float(v.double().sum()), float(v.numel())], dtype=torch.float64,
                        device=v.device if v.is_cuda else "cpu")
    _all_reduce(pair, group)
    return float(pair[0] / pair[1])
