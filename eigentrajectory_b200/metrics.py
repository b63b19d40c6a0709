"""ADE / FDE -- drop-in for ``utils/metrics.py:73-102`` running on libet_b200.so."""
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
