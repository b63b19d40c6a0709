"""ADE / FDE / TCC / COL -- drop-in for ``utils/metrics.py:30-155`` running on libet_b200.so."""
from __future__ import annotations

import weakref

import torch

from . import ops


_last = {"pred": None, "gt": None, "versions": None, "value": None}


def _scores(pred, gt):
    """(ADE, FDE, TCC) of ``(pred, gt)`` as numpy arrays from ONE kernel launch and ONE device-to-host copy.

    The reference's evaluation loop calls ``compute_batch_ade``, ``compute_batch_fde`` and ``compute_batch_tcc`` one
    after the other on the same tensors (utils/trainer.py:186-193); the last result is kept so that the second and
    third call of such a sequence cost nothing.  A hit needs the very same live tensor objects (weak references: a
    collected tensor can never match, even if a new one reuses its address) with unchanged autograd version counters
    (any in-place modification bumps them).  The entry holds only the small (3, N) host array, not the inputs."""
    hit = (_last["pred"] is not None and _last["pred"]() is pred and _last["gt"]() is gt
           and _last["versions"] == (pred._version, gt._version))
    if hit:
        return _last["value"]
    ade, fde, tcc = ops.ade_fde(pred, gt, want_tcc=True)
    host = torch.stack([ade, fde, tcc]).cpu().numpy()           # one D2H copy (and the only synchronisation)
    _last.update(pred=weakref.ref(pred), gt=weakref.ref(gt), versions=(pred._version, gt._version), value=host)
    return host


def compute_batch_ade_fde(pred, gt):
    r"""Both scores in ONE pass over pred (the reference recomputes the norm for each).

    Args:
        pred (torch.Tensor): (num_samples, num_ped, seq_len, 2)
        gt (torch.Tensor): (1, num_ped, seq_len, 2) or (num_ped, seq_len, 2)

    Returns:
        ADEs, FDEs (np.ndarray): (num_ped,) each
    """
    host = _scores(pred, gt)
    return host[0].copy(), host[1].copy()


def compute_batch_ade(pred, gt):
    r"""Compute ADE(average displacement error) scores for each pedestrian -> np.ndarray (num_ped,)"""
    return _scores(pred, gt)[0].copy()


def compute_batch_fde(pred, gt):
    r"""Compute FDE(final displacement error) scores for each pedestrian -> np.ndarray (num_ped,)"""
    return _scores(pred, gt)[1].copy()


def compute_batch_tcc(pred, gt):
    r"""Compute TCC(temporal correlation coefficient) scores for each pedestrian -> np.ndarray (num_ped,)
    (utils/metrics.py:105-130; same pass over pred as ADE/FDE)"""
    return _scores(pred, gt)[2].copy()


def compute_batch_col(pred, gt=None):
    r"""Compute COL(collision rate) scores for each pedestrian -> np.ndarray (num_ped,)
    (utils/metrics.py:133-155; ``gt`` is unused, as in the reference)"""
    return ops.col(pred).cpu().numpy()


def compute_batch_metric(pred, gt):
    r"""Get ADE, FDE, COL, TCC scores for each pedestrian (utils/metrics.py:30-70; returns tensors in the
    reference's order ADEs, FDEs, COLs, TCCs).  pred (S,N,T,2), gt (1,N,T,2) or (N,T,2)."""
    ade, fde, tcc = ops.ade_fde(pred, gt, want_tcc=True)
    return ade, fde, ops.col(pred), tcc
