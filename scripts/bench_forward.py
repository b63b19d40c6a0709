#!/usr/bin/env python
"""Per-scene EigenTrajectory.forward (+backward) latency: the launch-bound regime of config 4 (N = 2..57 pedestrians).

Times the model wrapper with a stub linear predictor behind the hook seam: fused glue (2 library launches, no
mask gathers) vs the reference's gather/scatter structure on the same kernels, and the reference arithmetic on the
CPU (oracle ops through the same wrapper structure is not available, so the CPU arm times the projection +
reconstruction + loss maths of model.py on the host)."""
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigentrajectory_b200 as et                                   # noqa: E402
from eigentrajectory_b200.synthetic import synthetic_trajectories   # noqa: E402

HP = dict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3, obs_svd=True, pred_svd=True)


class Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.W = torch.nn.Parameter(torch.randn(120, 8) * 0.1)

    def forward(self, x):
        return (self.W @ x).reshape(6, 20, -1).permute(0, 2, 1)


def main():
    dev = torch.device("cuda")
    hook = types.SimpleNamespace(model_forward_pre_hook=lambda C, o, info=None: torch.cat([C, o], dim=0),
                                 model_forward=lambda x, m: m(x), model_forward_post_hook=lambda y, info=None: y)
    obs_i, pred_i = synthetic_trajectories(20000, seed=5)
    slow = torch.arange(20000) % 3 == 0
    c = obs_i[:, -1:, :].clone()
    obs_i = torch.where(slow[:, None, None], c + (obs_i - c) * 0.3, obs_i)
    pred_i = torch.where(slow[:, None, None], c + (pred_i - c) * 0.3, pred_i)
    base = et.EigenTrajectory(Stub(), hook, et.DotDict(HP)).to(dev)
    base.calculate_parameters(obs_i.to(dev), pred_i.to(dev))
    sd = base.state_dict()
    rows = []
    for n in (8, 57, 512):
        obs, pred = obs_i[:n].to(dev), pred_i[:n].to(dev)
        for fused in (True, False):
            model = et.EigenTrajectory(Stub(), hook, et.DotDict(HP)).to(dev)
            model.load_state_dict(sd)
            model.fused = fused

            def step():
                out = model(obs, pred)
                (out["loss_eigentraj"] + out["loss_euclidean_ade"] + out["loss_euclidean_fde"]).backward()
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            l0 = et.launch_count()
            t0 = time.perf_counter()
            reps = 100
            for _ in range(reps):
                step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            rows.append({"n_peds": n, "fused": fused, "ms_per_forward_backward": 1e3 * dt,
                         "library_launches_per_step": (et.launch_count() - l0) / reps})
            print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "forward_latency.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
