"""Seeded synthetic pedestrian tracks for benchmarks and examples.

Generated on the CPU with a ``torch.Generator`` so that a CPU baseline and the GPU see identical
bits.  (The test-side checker keeps its own copy of this recipe; tests check that the two agree.)
"""
import math

import torch


def synthetic_trajectories(n, seed=0, t_obs=8, t_pred=12):
    """Constant-turn-rate walkers with velocity noise -> obs (n,t_obs,2), pred (n,t_pred,2), fp32 contiguous.

    p0 ~ U(-10,10)^2, heading ~ U(0,2pi), speed ~ U(0.2,0.8) m/frame (never static), turn rate ~ N(0,0.05^2),
    per-frame velocity noise N(0,0.03^2)."""
    g = torch.Generator().manual_seed(seed)
    T = t_obs + t_pred
    p0 = torch.rand(n, 1, 2, generator=g) * 20 - 10
    th0 = torch.rand(n, 1, generator=g) * (2 * math.pi)
    v = torch.rand(n, 1, generator=g) * 0.6 + 0.2
    om = torch.randn(n, 1, generator=g) * 0.05
    t = torch.arange(T, dtype=torch.float32)[None, :]
    ang = th0 + om * t
    vel = torch.stack([v * ang.cos(), v * ang.sin()], dim=-1) + torch.randn(n, T, 2, generator=g) * 0.03
    traj = (p0 + vel.cumsum(dim=1)).float()
    return traj[:, :t_obs].contiguous(), traj[:, t_obs:].contiguous()
