"""ETDescriptor -- drop-in for ``EigenTrajectory/descriptor.py:6-181`` running on libet_b200.so."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .normalizer import TrajNorm


class ETDescriptor(nn.Module):
    r"""EigenTrajectory descriptor model

    Args:
        hyper_params (DotDict): The hyper-parameters
        norm_ori (bool): Whether to normalize the trajectory with the origin
        norm_rot (bool): Whether to normalize the trajectory with the rotation
        norm_sca (bool): Whether to normalize the trajectory with the scale

    ``svd_method``: ``"auto"`` (default) = ``"gram"``: fp64 Gram pass + Jacobi eigen-solve, which on B200 is faster
    than the one-sided Jacobi at every size (95 vs 140-420 us for 16 columns, 155 vs 245-810 us for 24) and ~100x
    closer to the fp64 projector (1e-7 vs 1e-5; ``scripts/exp/svd_paths.py``); ``"jacobi"`` forces the shared-memory
    one-sided Jacobi kernel (``ops.svd_small``, whose strength is MANY small problems in one launch).
    """

    def __init__(self, hyper_params, norm_ori=True, norm_rot=True, norm_sca=True):
        super().__init__()

        self.hyper_params = hyper_params
        self.t_obs, self.t_pred = hyper_params.obs_len, hyper_params.pred_len
        self.obs_svd, self.pred_svd = hyper_params.obs_svd, hyper_params.pred_svd
        self.k = hyper_params.k
        self.s = hyper_params.num_samples
        self.dim = hyper_params.traj_dim
        assert self.dim == 2, "the B200 kernels are written for 2-D trajectories"
        self.traj_normalizer = TrajNorm(ori=norm_ori, rot=norm_rot, sca=norm_sca)
        self.svd_method = "auto"

        self.U_obs_trunc = nn.Parameter(torch.zeros((self.t_obs * self.dim, self.k)))
        self.U_pred_trunc = nn.Parameter(torch.zeros((self.t_pred * self.dim, self.k)))

    # ---- normaliser pass-throughs (descriptor.py:29-57) ----
    def normalize_trajectory(self, obs_traj, pred_traj=None):
        r"""Trajectory normalization -> (obs_traj_norm, pred_traj_norm | None)"""
        self.traj_normalizer.calculate_params(obs_traj)
        obs_traj_norm = self.traj_normalizer.normalize(obs_traj)
        pred_traj_norm = self.traj_normalizer.normalize(pred_traj) if pred_traj is not None else None
        return obs_traj_norm, pred_traj_norm

    def denormalize_trajectory(self, traj_norm):
        r"""Trajectory denormalization"""
        return self.traj_normalizer.denormalize(traj_norm)

    # ---- ET space (descriptor.py:59-89) ----
    def to_ET_space(self, traj, evec):
        r"""Transform Euclidean trajectories to EigenTrajectory coefficients: C (k,N) = evec^T M"""
        return ops.to_et_space(traj, evec)

    def to_Euclidean_space(self, C, evec):
        r"""Transform EigenTrajectory coefficients to Euclidean trajectories (N,T,dim)"""
        return ops.to_euclidean_space(C, evec, self.dim)

    # ---- eigen-basis (descriptor.py:91-114) ----
    def _basis(self, traj_norm, k):
        """(U (2T,k), S (k)) of an already normalised (N,T,2) tensor, on the compute device."""
        x = ops.to_dev(traj_norm)
        n, t = x.size(0), x.size(1)
        method = self.svd_method
        if method == "auto":
            method = "gram"
        if method == "jacobi":
            if not (n > 0 and ops.svd_small_fits(n, t)):
                raise ValueError(f"svd_method='jacobi': {n} rows x {2 * t} columns do not fit one SM's shared memory")
            U, S = ops.svd_small(x, k)
            return U[0], S[0]
        G, _ = ops.gram(x)
        return ops.eig_basis(G, k)

    def truncated_SVD(self, traj, k=None, full_matrices=False):
        r"""Truncated Singular Value Decomposition

        Returns (U_trunc (2T,k), S_trunc (k), V_trunc (N,k)) of the wide view M (2T,N).  Column signs
        are canonical (largest-magnitude component of each U column positive), which may differ from
        LAPACK's; V carries the matching sign (V = (U^T M)^T / S)."""
        assert traj.size(2) == self.dim  # NTC
        k = self.k if k is None else k
        U, S = self._basis(traj, k)
        C = ops.to_et_space(ops.to_dev(traj), U)          # (k,N)
        V = (C / S[:, None]).T
        return ops.back_to(U, traj), ops.back_to(S, traj), ops.back_to(V, traj)

    def parameter_initialization(self, obs_traj, pred_traj):
        r"""Initialize the ET descriptor parameters (for training only)

        Returns (pred_traj_norm, U_pred_trunc) for the anchor step.  Large inputs take ONE fused
        pass (normalise + both fp64 Gram matrices) followed by two tiny eigen-solves."""
        tn = self.traj_normalizer
        obs_d, pred_d = ops.to_dev(obs_traj), ops.to_dev(pred_traj)
        method = self.svd_method
        if method == "auto":
            method = "gram"
        if method == "jacobi":
            tn.calculate_params(obs_d)
            pred_norm_d = tn.normalize(pred_d)
            U_obs, _ = self._basis(tn.normalize(obs_d), self.k)
            U_pred, _ = self._basis(pred_norm_d, self.k)
        else:
            # one pass over the data: normaliser state, normalised futures and both Gram matrices; then one launch
            # for both eigen-solves
            G_obs, G_pred, pred_norm_d, state = ops.gram_init(obs_d, pred_d, tn.ori, tn.rot, tn.sca)
            for name, on, value in zip(("traj_ori", "traj_rot", "traj_sca"), (tn.ori, tn.rot, tn.sca), state):
                if on:
                    setattr(tn, name, value)
            (U_obs, _), (U_pred, _) = ops.eig_basis_pair(G_obs, G_pred, self.k)
        # state / outputs live where the caller's tensors live, as in the reference
        for name in ("traj_ori", "traj_rot", "traj_sca"):
            v = getattr(tn, name)
            if v is not None:
                setattr(tn, name, ops.back_to(v, obs_traj))
        self.U_obs_trunc = nn.Parameter(U_obs.to(self.U_obs_trunc.device))
        self.U_pred_trunc = nn.Parameter(U_pred.to(self.U_pred_trunc.device))
        return ops.back_to(pred_norm_d, pred_traj), ops.back_to(U_pred, pred_traj)

    # ---- per-batch path (descriptor.py:144-181) ----
    def projection(self, obs_traj, pred_traj=None):
        r"""Trajectory projection to the ET space -> (C_obs (k,N), C_pred (k,N) | None), detached"""
        tn = self.traj_normalizer
        C_obs, C_pred, (o, r, s) = ops.project(obs_traj, pred_traj, self.U_obs_trunc, self.U_pred_trunc,
                                               tn.ori, tn.rot, tn.sca)
        if tn.ori:
            tn.traj_ori = o
        if tn.rot:
            tn.traj_rot = r
        if tn.sca:
            tn.traj_sca = s
        return C_obs, C_pred

    def reconstruction(self, C_pred, anchor=None):
        r"""Trajectory reconstruction from the ET space: (k,N,S) -> (S,N,T,2), differentiable wrt C_pred.

        ``anchor`` (k,S) optionally fuses ETAnchor.forward (anchor.py:87) into the same kernel."""
        if C_pred.size(2) != self.s:
            C_pred = C_pred[:, :, :self.s]
        return ops.reconstruct(C_pred, self.U_pred_trunc, self.traj_normalizer.state(), anchor=anchor)

    def forward(self, C_pred):
        r"""Alias for reconstruction"""
        return self.reconstruction(C_pred)

    def project_reconstruct(self, obs_traj, pred_traj, want_coeffs=True, variant=0, out=None):
        r"""Rank-k round trip of (obs, pred) in one fused pass (the S=1 shape of
        script/descriptor_evaluation.py:94-107) -> (rec_obs, rec_pred, C_obs | None, C_pred | None)"""
        tn = self.traj_normalizer
        return ops.project_reconstruct(obs_traj, pred_traj, self.U_obs_trunc, self.U_pred_trunc, tn.ori, tn.rot,
                                       tn.sca, want_coeffs=want_coeffs, variant=variant, out=out)
