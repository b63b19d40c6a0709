"""EigenTrajectory wrapper -- same surface as ``EigenTrajectory/model.py:7-125``.

The caller of the hot path: it keeps the reference's constructor, ``calculate_parameters`` and
``forward`` (including the three predictor hook calls at model.py:93-95) so any ``baseline/``
predictor plugs in unchanged, and routes projection / anchor + reconstruction through the CUDA
kernels.  ``forward`` has two implementations with identical results: the default ``fused`` one
decides moving/static per pedestrian inside the kernels and computes the three training losses in one
more launch (three launches forward, one or two backward, no boolean-mask gathers and no host
synchronisation -- SURVEY.md section 8f-1); ``fused = False`` follows the reference's gather / scatter
structure group by group with the loss reductions as tensor algebra.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .anchor import ETAnchor
from .descriptor import ETDescriptor


class EigenTrajectory(nn.Module):
    r"""The EigenTrajectory model

    Args:
        baseline_model (nn.Module): The baseline model
        hook_func (dict): The bridge functions for the baseline model
        hyper_params (DotDict): The hyper-parameters
    """

    def __init__(self, baseline_model, hook_func, hyper_params):
        super().__init__()

        self.baseline_model = baseline_model
        self.hook_func = hook_func
        self.hyper_params = hyper_params
        self.t_obs, self.t_pred = hyper_params.obs_len, hyper_params.pred_len
        self.obs_svd, self.pred_svd = hyper_params.obs_svd, hyper_params.pred_svd
        self.k = hyper_params.k
        self.s = hyper_params.num_samples
        self.dim = hyper_params.traj_dim
        self.static_dist = hyper_params.static_dist

        self.ET_m_descriptor = ETDescriptor(hyper_params=hyper_params, norm_sca=True)
        self.ET_s_descriptor = ETDescriptor(hyper_params=hyper_params, norm_sca=False)
        self.ET_m_anchor = ETAnchor(hyper_params=hyper_params)
        self.ET_s_anchor = ETAnchor(hyper_params=hyper_params)
        self.fused = True

    def _can_fuse(self):
        tm, ts = self.ET_m_descriptor.traj_normalizer, self.ET_s_descriptor.traj_normalizer
        return (self.fused and (tm.ori, tm.rot, tm.sca) == (True, True, True) and (ts.ori, ts.rot, ts.sca) == (True, True, False))

    def _moving_mask(self, obs_traj):
        return (obs_traj[:, -1] - obs_traj[:, -3]).div(2).norm(p=2, dim=-1) > self.static_dist

    def calculate_parameters(self, obs_traj, pred_traj):
        r"""Calculate the ET descriptors of the EigenTrajectory model (model.py:34-56)

        Note:
            This function should be called once before training the model.
        """
        # Mask out static trajectory
        mask = self._moving_mask(obs_traj)
        obs_m_traj, pred_m_traj = obs_traj[mask], pred_traj[mask]
        obs_s_traj, pred_s_traj = obs_traj[~mask], pred_traj[~mask]

        # Descriptor initialization
        data_m = self.ET_m_descriptor.parameter_initialization(obs_m_traj, pred_m_traj)
        data_s = self.ET_s_descriptor.parameter_initialization(obs_s_traj, pred_s_traj)

        # Anchor generation
        self.ET_m_anchor.anchor_generation(*data_m)
        self.ET_s_anchor.anchor_generation(*data_s)

    def forward(self, obs_traj, pred_traj=None, addl_info=None):
        r"""The forward function of the EigenTrajectory model (model.py:58-125)

        Returns:
            output (dict): recon_traj (S,N,T,2) and, when pred_traj is given, the three loss scalars
        """
        if self._can_fuse():
            return self._forward_fused(obs_traj, pred_traj, addl_info)
        n_ped = obs_traj.size(0)
        dev = obs_traj.device

        # Filter out static trajectory
        mask = self._moving_mask(obs_traj)
        obs_m_traj, obs_s_traj = obs_traj[mask], obs_traj[~mask]
        pred_m_traj_gt = pred_traj[mask] if pred_traj is not None else None
        pred_s_traj_gt = pred_traj[~mask] if pred_traj is not None else None

        # Projection (fused normalise + U^T x; stores the normaliser state of each group)
        C_m_obs, C_m_pred_gt = self.ET_m_descriptor.projection(obs_m_traj, pred_m_traj_gt)
        C_s_obs, C_s_pred_gt = self.ET_s_descriptor.projection(obs_s_traj, pred_s_traj_gt)
        C_obs = torch.zeros((self.k, n_ped), dtype=torch.float, device=dev)
        C_obs[:, mask], C_obs[:, ~mask] = C_m_obs, C_s_obs  # KN

        # Absolute coordinate
        obs_m_ori = self.ET_m_descriptor.traj_normalizer.traj_ori.squeeze(dim=1).T
        obs_s_ori = self.ET_s_descriptor.traj_normalizer.traj_ori.squeeze(dim=1).T
        obs_ori = torch.zeros((2, n_ped), dtype=torch.float, device=dev)
        obs_ori[:, mask], obs_ori[:, ~mask] = obs_m_ori, obs_s_ori
        obs_ori -= obs_ori.mean(dim=1, keepdim=True)  # move scene to origin

        # Trajectory prediction (plugin seam, unchanged)
        input_data = self.hook_func.model_forward_pre_hook(C_obs, obs_ori, addl_info)
        output_data = self.hook_func.model_forward(input_data, self.baseline_model)
        C_pred_refine = self.hook_func.model_forward_post_hook(output_data, addl_info)

        # Anchor refinement + reconstruction in one kernel per group
        C_m_in, C_s_in = C_pred_refine[:, mask], C_pred_refine[:, ~mask]
        pred_m_traj_recon = self.ET_m_descriptor.reconstruction(C_m_in, anchor=self.ET_m_anchor.C_anchor)
        pred_s_traj_recon = self.ET_s_descriptor.reconstruction(C_s_in, anchor=self.ET_s_anchor.C_anchor)
        pred_traj_recon = torch.zeros((self.s, n_ped, self.t_pred, self.dim), dtype=torch.float, device=dev)
        pred_traj_recon[:, mask], pred_traj_recon[:, ~mask] = pred_m_traj_recon, pred_s_traj_recon

        output = {"recon_traj": pred_traj_recon}

        if pred_traj is not None:
            C_pred = torch.zeros((self.k, n_ped, self.s), dtype=torch.float, device=dev)
            C_pred[:, mask], C_pred[:, ~mask] = self.ET_m_anchor(C_m_in), self.ET_s_anchor(C_s_in)

            # Low-rank approximation for gt trajectory
            C_pred_gt = torch.zeros((self.k, n_ped), dtype=torch.float, device=dev)
            C_pred_gt[:, mask], C_pred_gt[:, ~mask] = C_m_pred_gt, C_s_pred_gt
            C_pred_gt = C_pred_gt.detach()

            # Loss calculation
            error_coefficient = (C_pred - C_pred_gt.unsqueeze(dim=-1)).norm(p=2, dim=0)
            error_displacement = (pred_traj_recon - pred_traj.unsqueeze(dim=0)).norm(p=2, dim=-1)
            output["loss_eigentraj"] = error_coefficient.min(dim=-1)[0].mean()
            output["loss_euclidean_ade"] = error_displacement.mean(dim=-1).min(dim=0)[0].mean()
            output["loss_euclidean_fde"] = error_displacement[:, :, -1].min(dim=0)[0].mean()

        return output

    def _forward_fused(self, obs_traj, pred_traj=None, addl_info=None):
        r"""Same results as ``forward`` with ``fused = False``; the moving/static decision, both projections, the
        anchor add and both reconstructions happen inside two kernels."""
        from . import ops
        dm, ds = self.ET_m_descriptor, self.ET_s_descriptor
        C_obs, C_pred_gt, state, moving = ops.forward_project(
            obs_traj, pred_traj, dm.U_obs_trunc, ds.U_obs_trunc, dm.U_pred_trunc, ds.U_pred_trunc, self.static_dist)
        C_obs, C_pred_gt = ops.back_to(C_obs, obs_traj), ops.back_to(C_pred_gt, obs_traj)

        # Absolute coordinate
        obs_ori = ops.back_to(state[0], obs_traj).squeeze(dim=1).T.clone()
        obs_ori -= obs_ori.mean(dim=1, keepdim=True)  # move scene to origin

        # Trajectory prediction (plugin seam, unchanged)
        input_data = self.hook_func.model_forward_pre_hook(C_obs, obs_ori, addl_info)
        output_data = self.hook_func.model_forward(input_data, self.baseline_model)
        C_pred_refine = self.hook_func.model_forward_post_hook(output_data, addl_info)
        if C_pred_refine.size(2) != self.s:
            C_pred_refine = C_pred_refine[:, :, :self.s]

        # Anchor refinement + reconstruction (+ the three losses of model.py:119-123 when training)
        am, an = self.ET_m_anchor.C_anchor.detach(), self.ET_s_anchor.C_anchor.detach()
        if pred_traj is None or obs_traj.size(0) == 0:
            recon = ops.forward_reconstruct(C_pred_refine, am, an, dm.U_pred_trunc, ds.U_pred_trunc, moving, state)
            output = {"recon_traj": recon}
            if pred_traj is not None:     # empty scene: the reference's means over zero pedestrians are NaN
                nan = recon.new_full((), float("nan"))
                output.update(loss_eigentraj=nan, loss_euclidean_ade=nan, loss_euclidean_fde=nan)
            return output
        recon, l_ec, l_ade, l_fde = ops.forward_reconstruct_losses(
            C_pred_refine, am, an, dm.U_pred_trunc, ds.U_pred_trunc, moving, state, C_pred_gt, pred_traj)
        output = {"recon_traj": recon, "loss_eigentraj": l_ec, "loss_euclidean_ade": l_ade, "loss_euclidean_fde": l_fde}

        return output
