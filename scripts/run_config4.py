#!/usr/bin/env python
"""BASELINE.json config 4: the reference's own evaluation loop (trainval.py:24-38 -> utils/trainer.py:172-195,
``ETSGCNTrainer.test``) on zara1 with the UNMODIFIED SGCN predictor (baseline/sgcn) behind the hook seam -- run twice
from the same ``state_dict``: once with the reference's EigenTrajectory / metrics modules (torch ops on the GPU) and
once with ``eigentrajectory_b200`` swapped in exactly as INTEGRATION.md section 1 describes (the ``model=`` class handed
to the trainer and the four ``compute_batch_*`` names the trainer imports).  Prints one JSON line with the
per-pedestrian ADE / FDE / TCC / COL differences and the time per scene of both runs.

TEST / MEASUREMENT INFRASTRUCTURE: needs the staged reference (oracle/make_ref.sh -> oracle/_ref) and a GPU.
"""
import contextlib
import io
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_trainer(ref_utils, ref_baseline, trainer_mod, model_cls, hp):
    hooks = ref_utils.DotDict({name: getattr(ref_baseline.sgcn, name)
                               for name in ("model_forward_pre_hook", "model_forward", "model_forward_post_hook")})
    args = types.SimpleNamespace(test=True, tag="config4", gpu_id="0", cfg="zara1")
    with contextlib.redirect_stdout(io.StringIO()):
        return trainer_mod.ETSGCNTrainer(base_model=ref_baseline.sgcn.TrajectoryPredictor, model=model_cls, hook_func=hooks,
                                         args=args, hyper_params=hp)


def run_test(trainer, repeats=2):
    """``trainer.test()`` (utils/trainer.py:172-195) -> per-pedestrian metric arrays and seconds per pass (best of ``repeats``)."""
    best, vals = None, None
    for _ in range(repeats):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stderr(io.StringIO()):          # tqdm
            means = trainer.test()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best = dt
        vals = {k: np.concatenate(m.data, axis=0) for k, m in trainer.stats_meter.items()}
    return vals, {k: float(v) for k, v in means.items()}, best


def main(scene="zara1", quiet=False):
    from oracle import ref_loader
    import eigentrajectory_b200 as et
    ref_ET, ref_utils, ref_baseline = ref_loader.load("EigenTrajectory", "utils", "baseline")
    import utils.trainer as trainer_mod                      # the reference's trainer module (staged copy)

    cfg = os.path.join(ref_loader.REF_ROOT, "config", "eigentrajectory-{baseline}-" + scene + ".json")
    hp = ref_utils.get_exp_config(cfg)
    hp.baseline = "sgcn"
    hp.dataset_dir = os.path.join(ref_loader.REF_ROOT, "datasets") + "/"
    hp.checkpoint_dir = "/tmp/et_config4_ckpt"

    # --- arm B first (ours): its descriptor initialisation is milliseconds on the GPU; the state_dict is then shared ---
    originals = {name: getattr(trainer_mod, name) for name in ("compute_batch_ade", "compute_batch_fde", "compute_batch_tcc",
                                                               "compute_batch_col")}
    for name in originals:
        setattr(trainer_mod, name, getattr(et, name))          # what editing utils/__init__.py's import line does
    ours = build_trainer(ref_utils, ref_baseline, trainer_mod, et.EigenTrajectory, hp)
    obs = torch.cat([ours.loader_train.dataset.obs_traj, ours.loader_val.dataset.obs_traj], dim=0)
    pred = torch.cat([ours.loader_train.dataset.pred_traj, ours.loader_val.dataset.pred_traj], dim=0)
    obs, pred = ref_utils.augment_trajectory(obs, pred)        # utils/trainer.py:47-52 (init_descriptor)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ours.model.calculate_parameters(obs.cuda(), pred.cuda())
    torch.cuda.synchronize()
    init_ours_s = time.perf_counter() - t0
    state = {k: v.detach().clone() for k, v in ours.model.state_dict().items()}
    anchors = {"moving": ours.model.ET_m_anchor.inertia_, "static": ours.model.ET_s_anchor.inertia_}
    launches = et.launch_count()
    vals_b, means_b, t_b = run_test(ours)
    launches = et.launch_count() - launches

    # --- arm A: the reference's own L2 modules (torch ops on the GPU), same predictor weights, same bases and anchors ---
    for name, fn in originals.items():
        setattr(trainer_mod, name, fn)
    ref = build_trainer(ref_utils, ref_baseline, trainer_mod, ref_ET.EigenTrajectory, hp)
    missing = ref.model.load_state_dict(state, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    vals_a, means_a, t_a = run_test(ref)

    scenes = len(ref.loader_test)
    peds = int(vals_a["ADE"].shape[0])
    out = {"config": "configs[3]: ET-SGCN evaluation loop, " + scene + " test split", "scenes": scenes, "pedestrians": peds,
           "init_descriptor_ours_ms": 1e3 * init_ours_s, "init_rows": int(obs.shape[0]),
           "ms_per_scene_reference_l2": 1e3 * t_a / scenes, "ms_per_scene_ours": 1e3 * t_b / scenes,
           "speedup": t_a / t_b, "library_launches_per_scene": launches / (2 * scenes), "anchor_inertia": anchors}
    for k in ("ADE", "FDE", "TCC", "COL"):
        a, b = vals_a[k], vals_b[k]
        assert a.shape == b.shape
        out[k] = {"mean_reference": means_a[k], "mean_ours": means_b[k], "pedestrians_differing": int((a != b).sum()),
                  "max_abs_diff": float(np.abs(a - b).max()), "max_rel_diff": float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-30))}
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    main(*(sys.argv[1:2]))
