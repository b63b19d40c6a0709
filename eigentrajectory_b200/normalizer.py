"""TrajNorm -- the per-pedestrian trajectory normaliser, API-compatible with the reference's
``EigenTrajectory/normalizer.py:4-62`` and computed by ``libet_b200.so``.

State of a batch of N pedestrians, derived from the observed frames ``obs (N, T_obs, 2)``:
  * ``traj_ori (N,1,2)``  the last observed position (translation),
  * ``traj_rot (N,2,2)``  ``[[c,-s],[s,c]]`` with ``(c, s)`` the heading of ``obs[-1] - obs[-3]`` (rotation),
  * ``traj_sca (N,1,1)``  ``2 / |obs[-1] - obs[-3]|`` (scale; ``inf`` for a pedestrian that did not move, as in the
    reference).
``normalize`` maps ``x -> ((x - ori) @ rot) * sca``, ``denormalize`` inverts it; each stage is optional.

The heading is taken as ``d/|d|`` instead of ``cos/sin(atan2(d))`` (<= 3e-7 absolute difference; the zero vector gives
the identity rotation exactly as ``atan2(0, 0) = 0`` does).
"""
from __future__ import annotations

from . import ops

_STATE = ("traj_ori", "traj_rot", "traj_sca")


def _state_property(name):
    slot = "_" + name

    def get(self):
        self._materialise()
        return getattr(self, slot)

    def put(self, value):
        self._materialise()            # an explicit assignment to one field must not be overwritten later
        setattr(self, slot, value)

    return property(get, put)


class TrajNorm:
    """Translation / rotation / scale normaliser for trajectories of shape ``(num_peds, length_of_time, 2)``.

    ``ori``, ``rot``, ``sca`` switch the three stages on or off (all on by default).

    ``traj_ori`` / ``traj_rot`` / ``traj_sca`` are plain public attributes for the caller.  Internally they may be held
    as a *deferred row selection* (:meth:`set_deferred`): the fused ``EigenTrajectory.forward`` computes the state of
    every pedestrian of the scene in one kernel and hands each group's normaliser (full state, row mask) -- the
    boolean gather, which needs a host synchronisation, only happens if somebody actually reads the state."""

    traj_ori, traj_rot, traj_sca = (_state_property(n) for n in _STATE)

    def __init__(self, ori=True, rot=True, sca=True):
        self.ori, self.rot, self.sca = ori, rot, sca
        self._deferred = None
        for name in _STATE:
            setattr(self, "_" + name, None)

    def set_deferred(self, full_state, rows):
        """State of this normaliser := ``full_state[i][rows]`` for the enabled stages, gathered on first use."""
        self._deferred = (full_state, rows)

    def _materialise(self):
        if self._deferred is None:
            return
        (full, rows), self._deferred = self._deferred, None
        for name, on, value in zip(_STATE, self._enabled(), full):
            if on and value is not None:
                setattr(self, "_" + name, value[rows])

    def _enabled(self):
        return self.ori, self.rot, self.sca

    def calculate_params(self, traj):
        """Derive the state of this batch from its observed frames (normalizer.py:17-28); stages that are switched
        off keep whatever state they had."""
        fresh = ops.norm_params(traj, *self._enabled())
        for name, on, value in zip(_STATE, self._enabled(), fresh):
            if on:
                setattr(self, name, value)

    def get_params(self):
        """``(ori, rot, sca, traj_ori, traj_rot, traj_sca)``: flags and state, for hand-over to another instance."""
        return (*self._enabled(), *(getattr(self, name) for name in _STATE))

    def set_params(self, ori, rot, sca, traj_ori, traj_rot, traj_sca):
        """Adopt flags and state obtained from :meth:`get_params`."""
        self.ori, self.rot, self.sca = ori, rot, sca
        self.traj_ori, self.traj_rot, self.traj_sca = traj_ori, traj_rot, traj_sca

    def state(self):
        """``(ori, rot, sca)`` tensors with disabled stages as ``None`` -- the form the fused kernels take."""
        return tuple(getattr(self, name) if on else None for name, on in zip(_STATE, self._enabled()))

    def normalize(self, traj):
        """``((traj - ori) @ rot) * sca`` over the enabled stages (normalizer.py:42-51)."""
        return ops.normalize(traj, *self.state()) if any(self._enabled()) else traj

    def denormalize(self, traj):
        """``((traj / sca) @ rot^T) + ori`` over the enabled stages (normalizer.py:53-62)."""
        return ops.denormalize(traj, *self.state()) if any(self._enabled()) else traj
