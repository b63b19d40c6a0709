"""ADE / FDE / TCC / COL -- drop-in for ``utils/metrics.py:30-155`` running on libet_b200.so."""
from __future__ import annotations

from . import ops


def compute_batch_ade_fde(pred, gt):
    r"""Both scores in ONE pass over pred (the reference recomputes the norm for each).

    Args:
        pred (torch.Tensor): (num_samples, num_ped, seq_len, 2)
        gt (torch.Tensor): (1, num_ped, seq_len, 2) or (num_ped, seq_len, 2)

    Returns:
        ADEs, FDEs (np.ndarray): (num_ped,) each
    """
    ade, fde = ops.ade_fde(pred, gt)
    return ade.cpu().numpy(), fde.cpu().numpy()


def compute_batch_ade(pred, gt):
    r"""Compute ADE(average displacement error) scores for each pedestrian -> np.ndarray (num_ped,)"""
    return ops.ade_fde(pred, gt)[0].cpu().numpy()


def compute_batch_fde(pred, gt):
    r"""Compute FDE(final displacement error) scores for each pedestrian -> np.ndarray (num_ped,)"""
    return ops.ade_fde(pred, gt)[1].cpu().numpy()


def compute_batch_tcc(pred, gt):
    r"""Compute TCC(temporal correlation coefficient) scores for each pedestrian -> np.ndarray (num_ped,)
    (utils/metrics.py:105-130; same pass over pred as ADE/FDE)"""
    return ops.ade_fde(pred, gt, want_tcc=True)[2].cpu().numpy()


def compute_batch_col(pred, gt=None):
    r"""Compute COL(collision rate) scores for each pedestrian -> np.ndarray (num_ped,)
    (utils/metrics.py:133-155; ``gt`` is unused, as in the reference)"""
    return ops.col(pred).cpu().numpy()


def compute_batch_metric(pred, gt):
    r"""Get ADE, FDE, COL, TCC scores for each pedestrian (utils/metrics.py:30-70; returns tensors in the
    reference's order ADEs, FDEs, COLs, TCCs).  pred (S,N,T,2), gt (1,N,T,2) or (N,T,2)."""
    ade, fde, tcc = ops.ade_fde(pred, gt, want_tcc=True)
    return ade, fde, ops.col(pred), tcc
