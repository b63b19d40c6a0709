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

import torch.nn as nn

from .anchor import ETAnchor
from .descriptor import ETDescriptor


class EigenTrajectory(nn.Module):
    """Predictor-agnostic EigenTrajectory wrapper (same constructor and ``state_dict`` keys as the reference's class,
    model.py:7-32).

    ``baseline_model``: the predictor network behind the hook seam.  ``hook_func``: its three bridge functions
    (``model_forward_pre_hook``, ``model_forward``, ``model_forward_post_hook``).  ``hyper_params``: DotDict with
    ``obs_len, pred_len, k, num_samples, traj_dim, static_dist, obs_svd, pred_svd``.

    Two descriptor / anchor pairs are kept: ``ET_m_*`` for moving pedestrians (scale normalisation on) and ``ET_s_*``
    for static ones (scale off); ``static_dist`` decides the group per pedestrian.
    """

    def __init__(self, baseline_model, hook_func, hyper_params):
        super().__init__()
        hp = hyper_params
        self.baseline_model, self.hook_func, self.hyper_params = baseline_model, hook_func, hp
        self.t_obs, self.t_pred = hp.obs_len, hp.pred_len
        self.obs_svd, self.pred_svd = hp.obs_svd, hp.pred_svd
        self.k, self.s, self.dim, self.static_dist = hp.k, hp.num_samples, hp.traj_dim, hp.static_dist
        # registration order fixes the state_dict key order: m-descriptor, s-descriptor, m-anchor, s-anchor
        self.ET_m_descriptor = ETDescriptor(hyper_params=hp, norm_sca=True)
        self.ET_s_descriptor = ETDescriptor(hyper_params=hp, norm_sca=False)
        self.ET_m_anchor = ETAnchor(hyper_params=hp)
        self.ET_s_anchor = ETAnchor(hyper_params=hp)
        self.fused = True          # mask-free forward (two or three launches); False: the reference's gather / scatter shape

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
        moving = self._moving_mask(obs_traj)
        for rows, descriptor, anchor in ((moving, self.ET_m_descriptor, self.ET_m_anchor),
                                         (~moving, self.ET_s_descriptor, self.ET_s_anchor)):
            # eigen-bases of the group, then its anchors from the normalised futures in that basis
            pred_norm, U_pred = descriptor.parameter_initialization(obs_traj[rows], pred_traj[rows])
            anchor.anchor_generation(pred_norm, U_pred)

    def forward(self, obs_traj, pred_traj=None, addl_info=None):
        r"""The forward function of the EigenTrajectory model (model.py:58-125)

        Returns:
            output (dict): recon_traj (S,N,T,2) and, when pred_traj is given, the three loss scalars
        """
        if self._can_fuse():
            return self._forward_fused(obs_traj, pred_traj, addl_info)
        return self._forward_grouped(obs_traj, pred_traj, addl_info)

    @staticmethod
    def _merge(moving, part_m, part_s, shape, axis):
        """Tensor of ``shape`` whose slices along ``axis`` come from ``part_m`` where ``moving`` is set and from
        ``part_s`` elsewhere (the scatter the reference writes out for every quantity, model.py:84-117)."""
        full = part_m.new_zeros(shape)
        sel = [slice(None)] * len(shape)
        sel[axis] = moving
        full[tuple(sel)] = part_m
        sel[axis] = ~moving
        full[tuple(sel)] = part_s
        return full

    def _forward_grouped(self, obs_traj, pred_traj, addl_info):
        """``forward`` in the reference's group-by-group structure: gather the moving / static pedestrians, run each
        group through its own descriptor and anchor, scatter back; the losses are tensor algebra (model.py:73-123)."""
        n = obs_traj.size(0)
        moving = self._moving_mask(obs_traj)
        static = ~moving
        have_gt = pred_traj is not None
        desc = {True: self.ET_m_descriptor, False: self.ET_s_descriptor}
        anchor = {True: self.ET_m_anchor, False: self.ET_s_anchor}

        # projection of each group (fused normalise + U^T x; each descriptor keeps its group's normaliser state)
        coef_obs, coef_gt, origin = {}, {}, {}
        for grp, rows in ((True, moving), (False, static)):
            coef_obs[grp], coef_gt[grp] = desc[grp].projection(obs_traj[rows], pred_traj[rows] if have_gt else None)
            origin[grp] = desc[grp].traj_normalizer.traj_ori.squeeze(dim=1).T
        C_obs = self._merge(moving, coef_obs[True], coef_obs[False], (self.k, n), 1)              # (k, N)
        obs_ori = self._merge(moving, origin[True], origin[False], (2, n), 1)
        obs_ori -= obs_ori.mean(dim=1, keepdim=True)                                             # scene-centred

        # predictor plugin: the three bridge calls of model.py:93-95, untouched
        hooks = self.hook_func
        C_refined = hooks.model_forward_post_hook(
            hooks.model_forward(hooks.model_forward_pre_hook(C_obs, obs_ori, addl_info), self.baseline_model), addl_info)

        # anchor refinement + reconstruction, one kernel per group
        refined = {True: C_refined[:, moving], False: C_refined[:, static]}
        recon = {g: desc[g].reconstruction(refined[g], anchor=anchor[g].C_anchor) for g in (True, False)}
        recon_traj = self._merge(moving, recon[True], recon[False], (self.s, n, self.t_pred, self.dim), 1)
        output = {"recon_traj": recon_traj}
        if not have_gt:
            return output

        # losses (model.py:119-123): coefficient error, ADE and FDE, each min over the S samples then mean over N
        C_pred = self._merge(moving, anchor[True](refined[True]), anchor[False](refined[False]), (self.k, n, self.s), 1)
        C_gt = self._merge(moving, coef_gt[True], coef_gt[False], (self.k, n), 1).detach()
        err_coef = (C_pred - C_gt.unsqueeze(dim=-1)).norm(p=2, dim=0)                             # (N, S)
        err_disp = (recon_traj - pred_traj.unsqueeze(dim=0)).norm(p=2, dim=-1)                    # (S, N, T)
        output["loss_eigentraj"] = err_coef.min(dim=-1)[0].mean()
        output["loss_euclidean_ade"] = err_disp.mean(dim=-1).min(dim=0)[0].mean()
        output["loss_euclidean_fde"] = err_disp[:, :, -1].min(dim=0)[0].mean()
        return output

    def _forward_fused(self, obs_traj, pred_traj=None, addl_info=None):
        r"""Same results as ``forward`` with ``fused = False``; the moving/static decision, both projections, the
        anchor add and both reconstructions happen inside two kernels."""
        from . import ops
        dm, ds = self.ET_m_descriptor, self.ET_s_descriptor
        C_obs, C_pred_gt, state, moving = ops.forward_project(
            obs_traj, pred_traj, dm.U_obs_trunc, ds.U_obs_trunc, dm.U_pred_trunc, ds.U_pred_trunc, self.static_dist)
        C_obs, C_pred_gt = ops.back_to(C_obs, obs_traj), ops.back_to(C_pred_gt, obs_traj)
        # each group's normaliser sees its rows of the per-scene state exactly as after the reference's projection()
        # calls (model.py:80-81) -- as a deferred selection, so no boolean gather happens unless the state is read
        dm.traj_normalizer.set_deferred(state, moving)
        ds.traj_normalizer.set_deferred(state, ~moving)

        # predictor input: last observed positions relative to the centre of the scene (model.py:88-90)
        obs_ori = ops.back_to(state[0], obs_traj).squeeze(dim=1).T.clone()
        obs_ori -= obs_ori.mean(dim=1, keepdim=True)

        # the three bridge calls of the plugin seam, as the reference makes them (model.py:93-95)
        hooks = self.hook_func
        C_pred_refine = hooks.model_forward_post_hook(
            hooks.model_forward(hooks.model_forward_pre_hook(C_obs, obs_ori, addl_info), self.baseline_model), addl_info)
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
