"""TrajNorm -- drop-in for ``EigenTrajectory/normalizer.py:4-62`` running on libet_b200.so."""
from __future__ import annotations

from . import ops


class TrajNorm:
    r"""Normalize trajectory with shape (num_peds, length_of_time, 2)

    Args:
        ori (bool): Whether to normalize the trajectory with the origin
        rot (bool): Whether to normalize the trajectory with the rotation
        sca (bool): Whether to normalize the trajectory with the scale

    Same public surface as the reference (``ori/rot/sca`` flags and the stored per-batch state
    ``traj_ori (N,1,2)``, ``traj_rot (N,2,2)``, ``traj_sca (N,1,1)``).  The rotation is obtained as
    ``d/|d|`` instead of ``cos/sin(atan2(d))`` (<= 3e-7 absolute difference; the zero vector maps
    to the identity exactly as ``atan2(0, 0) = 0`` does).
    """

    def __init__(self, ori=True, rot=True, sca=True):
        self.ori, self.rot, self.sca = ori, rot, sca
        self.traj_ori, self.traj_rot, self.traj_sca = None, None, None

    def calculate_params(self, traj):
        r"""Calculate the normalization parameters (normalizer.py:17-28)"""
        o, r, s = ops.norm_params(traj, self.ori, self.rot, self.sca)
        if self.ori:
            self.traj_ori = o
        if self.rot:
            self.traj_rot = r
        if self.sca:
            self.traj_sca = s

    def get_params(self):
        r"""Get the normalization parameters"""
        return self.ori, self.rot, self.sca, self.traj_ori, self.traj_rot, self.traj_sca

    def set_params(self, ori, rot, sca, traj_ori, traj_rot, traj_sca):
        r"""Set the normalization parameters"""
        self.ori, self.rot, self.sca = ori, rot, sca
        self.traj_ori, self.traj_rot, self.traj_sca = traj_ori, traj_rot, traj_sca

    def state(self):
        """(ori, rot, sca) with disabled stages as None -- what the fused kernels take."""
        return (self.traj_ori if self.ori else None, self.traj_rot if self.rot else None,
                self.traj_sca if self.sca else None)

    def normalize(self, traj):
        r"""Normalize the trajectory (normalizer.py:42-51)"""
        if not (self.ori or self.rot or self.sca):
            return traj
        return ops.normalize(traj, *self.state())

    def denormalize(self, traj):
        r"""Denormalize the trajectory (normalizer.py:53-62)"""
        if not (self.ori or self.rot or self.sca):
            return traj
        return ops.denormalize(traj, *self.state())
