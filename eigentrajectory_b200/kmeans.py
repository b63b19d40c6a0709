"""BatchKMeans -- drop-in for ``EigenTrajectory/kmeans.py:7-272`` running on libet_b200.so."""
from __future__ import annotations

from time import time

import numpy as np
import torch
import torch.nn as nn

from . import ops


def _as_ldn(t):
    """(..., d, N) -> contiguous (l, d, N) fp32 on the compute device, plus the leading shape."""
    lead = tuple(t.shape[:-2])
    x = ops.to_dev(t).reshape(-1, t.size(-2), t.size(-1))
    return x, lead


class BatchKMeans(nn.Module):
    """A batch of independent k-means problems clustered side by side (the reference's class of the same name,
    kmeans.py:7-272): data ``(..., d, N)``, centroids ``(..., d, K)``, one problem per leading index.

    ``n_clusters``: K.  ``n_redo``: restarts from different seeds, the fit with the lowest inertia wins (default 1).
    ``max_iter`` (100) and ``tol`` (1e-4): Lloyd stops after ``max_iter`` updates or once the squared centroid shift,
    summed over the whole batch, is ``<= tol``.  ``init_mode``: ``'kmeans++'`` (default) = the reference's
    deterministic farthest-point rule from a random first point, ``'random'`` = K distinct random points, ``'d2'`` (an
    addition) = k-means++ proper, D^2 sampling with greedy local trials as sklearn seeds the reference's anchors.

    ``fused`` (default): the whole Lloyd loop of ``fit`` runs in ONE persistent cooperative kernel
    (``et_kmeans_lloyd``: grid barriers instead of relaunches, convergence test on the device).
    ``fused = False`` / ``verbose``: one assign + finalize launch pair per iteration, ``sync_every``
    iterations enqueued between host checks of the device-side convergence flag (the reference
    synchronises every iteration at kmeans.py:239; the stopping iteration is the same, later
    launches are no-ops).  Both produce identical centroids, labels and iteration counts.
    """

    def __init__(self, n_clusters, n_redo=1, max_iter=100, tol=1e-4, init_mode="kmeans++", verbose=False):
        super(BatchKMeans, self).__init__()
        self.n_redo = n_redo
        self.n_clusters = n_clusters
        self.max_iter = max_iter
        self.tol = tol
        self.init_mode = init_mode
        self.verbose = verbose
        self.sync_every = 8
        self.fused = True        # whole Lloyd loop in one persistent kernel; False: one launch pair per iteration
        self.n_iter_ = None
        self.inertia_ = None

        self.register_buffer("centroids", None)

    def load_state_dict(self, state_dict, **kwargs):
        """``centroids`` is a buffer that does not exist before the first fit, so a plain ``load_state_dict`` would
        reject it: top-level entries are (re-)registered as buffers, dotted ones go to the child modules
        (kmeans.py:32-43)."""
        own = {key: value for key, value in state_dict.items() if "." not in key}
        for key, value in own.items():
            assert hasattr(self, key), f"attribute {key} does not exist"
            delattr(self, key)
            self.register_buffer(key, value)
        for child_name, child in self.named_children():
            prefix = child_name + "."
            child.load_state_dict({key[len(prefix):]: value for key, value in state_dict.items() if key.startswith(prefix)})

    # ---- small static helpers: tensor algebra on (l,d,K)-sized or already reduced data ----
    @staticmethod
    def calculate_error(a, b):
        """Squared Euclidean distance between two centroid sets, summed over everything (kmeans.py:45-51)."""
        return (a - b).square_().sum()

    @staticmethod
    def calculate_inertia(a):
        """Inertia from best similarities: similarities are negative squared distances (kmeans.py:53-57)."""
        return (-a).mean()

    @staticmethod
    def euc_sim(a, b):
        r"""Batched negative squared Euclidean distance, (..., d, m) x (..., d, n) -> (..., m, n).

        Utility kept for API parity (kmeans.py:59-76); it materialises the full matrix with tensor
        algebra on whatever device its inputs live.  The clustering itself never calls it: get_labels
        fuses similarity, arg-max and the centroid accumulation in one CUDA kernel."""
        sq_a, sq_b = a.pow(2).sum(dim=-2), b.pow(2).sum(dim=-2)
        sim = torch.matmul(a.transpose(-2, -1), b)
        sim.mul_(2).sub_(sq_a.unsqueeze(-1)).sub_(sq_b.unsqueeze(-2))      # same in-place order as the reference
        return sim

    # ---- seeding (kmeans.py:78-141) ----
    def kmeanspp(self, data):
        r"""Initialize centroids with the reference's farthest-point 'k-means++' (..., d, N) -> (..., d, K)"""
        x, lead = _as_ldn(data)
        first = np.random.randint(x.size(-1))
        cent = ops.kmeans_farthest_init(x, self.n_clusters, first)
        return ops.back_to(cent.reshape(*lead, x.size(1), self.n_clusters), data)

    def d2_sampling(self, data, n_local_trials=None):
        r"""k-means++ seeding proper (D^2 sampling with greedy local trials, as ``sklearn.cluster.KMeans(init="k-means++")``
        does for the reference's anchors, anchor.py:65-71) -- ``init_mode = "d2"``, an addition to the reference's two
        modes.  The random numbers come from NumPy's global generator, like the reference's other modes."""
        x, lead = _as_ldn(data)
        trials = n_local_trials or (2 + int(np.log(self.n_clusters)))
        uniform = np.random.random_sample((x.size(0), self.n_clusters, trials))
        cent = ops.kmeans_d2_init(x, self.n_clusters, uniform, trials)
        return ops.back_to(cent.reshape(*lead, x.size(1), self.n_clusters), data)

    def initialize_centroids(self, data):
        """Starting centroids ``(..., d, K)`` according to ``init_mode`` (kmeans.py:114-141)."""
        if self.init_mode == "d2":
            start = self.d2_sampling(data)
            how = "with D^2 sampling (k-means++)"
        elif self.init_mode == "kmeans++":
            start = self.kmeanspp(data).clone()
            how = "with kmeans++"
        elif self.init_mode == "random":
            pick = np.random.choice(data.size(-1), size=[self.n_clusters], replace=False)
            start = data[..., torch.as_tensor(pick, device=data.device)].clone()
            how = "randomly"
        else:
            raise NotImplementedError
        if self.verbose:
            print(f"centroids are initialized {how}.")
        return start

    # ---- one Lloyd half-step each (kmeans.py:143-198) ----
    def get_labels(self, data, centroids):
        r"""Compute labels of data -> (maxsims (..., N) fp32, labels (..., N) int64)"""
        x, lead = _as_ldn(data)
        c, _ = _as_ldn(centroids)
        c = c.to(x.device)
        maxsims, labels = ops.kmeans_assign(x, c)
        n = x.size(-1)
        return (ops.back_to(maxsims.reshape(*lead, n), data), ops.back_to(labels.reshape(*lead, n), data))

    def compute_centroids_loop(self, data, labels):
        r"""Compute centroids of data given labels (kmeans.py:160-184): masked sum / count per cluster
        (fp64 accumulation; an empty cluster yields NaN exactly as the reference's 0/0 does)."""
        x, lead = _as_ldn(data)
        l, d, n = x.shape
        lab = labels.reshape(l, n).to(device=x.device, dtype=torch.int64).contiguous()
        acc = ops.KMeansWorkspace(l, d, self.n_clusters, x.device)
        ops.kmeans_accumulate(x, lab, acc)
        cent = torch.empty((l, d, self.n_clusters), device=x.device)
        ops.kmeans_finalize(acc, None, cent)
        return ops.back_to(cent.reshape(*lead, d, self.n_clusters), data)

    def compute_centroids(self, data, labels):
        r"""Compute centroids of data"""
        return self.compute_centroids_loop(data, labels)

    # ---- fit / predict (kmeans.py:200-272) ----
    def _lloyd(self, x, centroids, acc):
        """Run Lloyd iterations from ``centroids`` on the device.  Returns (labels, centroids, n_iter, error, inertia)."""
        l, d, n = x.shape
        if self.fused and not self.verbose and self.max_iter >= 1 and l <= ops.KMEANS_FUSED_MAX_BATCH:
            # one persistent cooperative launch for the whole loop (no host round trip per iteration); the single
            # host read below is the result the caller needs anyway
            labels, final = ops.kmeans_lloyd(x, centroids.contiguous(), acc, self.max_iter, self.tol)
            _, n_iter = (int(v) for v in acc.status.tolist())
            inertia = -(acc.simsum_last / n).mean()
            return labels, final, n_iter, acc.err.clone(), inertia
        bufs = [centroids.contiguous().clone(), torch.empty_like(centroids)]
        acc.reset()
        done = 0
        n_iter = 0
        while done < self.max_iter:
            chunk = min(self.sync_every, self.max_iter - done)
            for j in range(done, done + chunk):
                cur, nxt = bufs[j % 2], bufs[(j + 1) % 2]
                ops.kmeans_assign(x, cur, want_labels=False, want_maxsims=False, acc=acc, simsum=acc.simsum,
                                  status=acc.status)
                ops.kmeans_finalize(acc, cur, nxt, tol=self.tol, use_status=True)
            done += chunk
            converged, n_iter = (int(v) for v in acc.status.tolist())    # the only host sync per chunk
            if self.verbose:
                print(f"----{n_iter} iterations done, converged={bool(converged)}")
            if converged:
                break
        final = bufs[n_iter % 2]
        before = bufs[(n_iter - 1) % 2]
        _, labels = ops.kmeans_assign(x, before, want_labels=True, want_maxsims=False)
        inertia = -(acc.simsum_last / n).mean()
        return labels, final.clone(), n_iter, acc.err.clone(), inertia

    def _workspace(self, l, d, device):
        """Device scratch of a fit, kept on the object and reused while the problem shape stays the same (a fit of this
        size allocates ~4 MB of zeroed scratch otherwise; every launch that needs clean counters zeroes them itself)."""
        key = (l, d, self.n_clusters, device, torch.cuda.current_stream(device).cuda_stream)
        if getattr(self, "_acc_key", None) != key:
            self._acc, self._acc_key = ops.KMeansWorkspace(l, d, self.n_clusters, device), key
        return self._acc

    def fit(self, data, centroids=None):
        """Cluster ``data (l, d, N)`` and return the labels ``(l, N)`` of the best restart; the winning centroids
        are kept in the ``centroids`` buffer.  ``centroids (l, d, K)``, if given, seed the first restart
        (kmeans.py:200-259)."""
        assert data.is_contiguous(), "use .contiguous()"
        x, lead = _as_ldn(data)
        l, d, n = x.shape
        acc = self._workspace(l, d, x.device)

        best_centroids = best_labels = None
        best_inertia = 1e32
        if self.verbose:
            tm = time()
        for i in range(self.n_redo):
            if self.verbose:
                tm_i = time()
            if centroids is None:
                c0 = self.initialize_centroids(x)
            else:
                c0, _ = _as_ldn(centroids)
                c0 = c0.to(x.device)
            labels, cent, n_iter, error, inertia = self._lloyd(x, c0, acc)
            inertia = float(inertia)
            if inertia < best_inertia:
                best_centroids, best_labels, best_inertia = cent, labels, inertia
                self.n_iter_ = n_iter
            centroids = None
            if self.verbose:
                print(f"--{i}th redo finished, error: {float(error)}, inertia: {inertia}, "
                      f"time spent:{round(time() - tm_i, 4)} sec")
        if best_centroids is None:          # every redo produced NaN inertia: keep the last, as a max() would not
            best_centroids, best_labels = cent, labels
        self.inertia_ = best_inertia
        k = self.n_clusters
        self.register_buffer("centroids", ops.back_to(best_centroids.reshape(*lead, d, k), data))
        if self.verbose:
            print(f"finished {self.n_redo} redos in {round(time() - tm, 4)} sec, final_inertia: {best_inertia}")
        return ops.back_to(best_labels.reshape(*lead, n), data)

    def predict(self, query):
        r"""Predict the closest cluster center each sample in query belongs to."""
        _, labels = self.get_labels(query, self.centroids)
        return labels
