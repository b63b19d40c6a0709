"""ETAnchor -- drop-in for ``EigenTrajectory/anchor.py:5-88`` running on libet_b200.so."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .kmeans import BatchKMeans


class ETAnchor(nn.Module):
    r"""EigenTrajectory anchor model

    Args:
        hyper_params (DotDict): The hyper-parameters

    ``anchor_generation`` clusters the pred coefficients with the GPU ``BatchKMeans`` (Lloyd with the
    reference's deterministic farthest-point seeding, ``n_redo`` restarts from different first
    points).  The reference calls ``sklearn.cluster.KMeans(n_init=10, random_state=0)`` here
    (anchor.py:65-71); that third-party result is not bit-reproducible by construction, so anchor
    parity is statistical (inertia), not bitwise -- see DESIGN.md.
    """

    def __init__(self, hyper_params):
        super().__init__()

        self.hyper_params = hyper_params
        self.k = hyper_params.k
        self.s = hyper_params.num_samples
        self.dim = hyper_params.traj_dim
        self.n_redo = 10          # mirrors sklearn's n_init=10
        self.kmeans_seed = 0      # mirrors random_state=0

        self.C_anchor = nn.Parameter(torch.zeros((self.k, self.s)))

    def to_ET_space(self, traj, evec):
        r"""Transform Euclidean trajectories to EigenTrajectory coefficients (anchor.py:22-36)"""
        return ops.to_et_space(traj, evec)

    def to_Euclidean_space(self, C, evec):
        r"""Transform EigenTrajectory coefficients to Euclidean trajectories (anchor.py:38-52)"""
        return ops.to_euclidean_space(C, evec, self.dim)

    def anchor_generation(self, pred_traj_norm, U_pred_trunc):
        r"""Anchor generation on EigenTrajectory space (anchor.py:54-74)

        Note:
            This function should be called once before training the model.
        """
        # Trajectory projection: (k, N) on the GPU, already the (l=1, d=k, N) layout k-means wants
        C_pred = ops.to_et_space(ops.to_dev(pred_traj_norm), ops.to_dev(U_pred_trunc)).unsqueeze(0)

        km = BatchKMeans(n_clusters=self.s, n_redo=self.n_redo)
        rng_state = np.random.get_state()
        np.random.seed(self.kmeans_seed)
        try:
            km.fit(C_pred)
        finally:
            np.random.set_state(rng_state)
        C_anchor = km.centroids[0]                      # (k, s)

        # Register anchors as model parameters
        self.C_anchor = nn.Parameter(C_anchor.to(self.C_anchor.device))

    def forward(self, C_pred):
        r"""Anchor refinement on EigenTrajectory space: C_anchor[:, None, :] + C_pred (anchor.py:76-88).

        A broadcast add kept in autograd; ``ETDescriptor.reconstruction(C, anchor=...)`` fuses the same
        add into the reconstruction kernel and is what ``EigenTrajectory.forward`` uses."""
        return self.C_anchor.unsqueeze(dim=1).detach() + C_pred
