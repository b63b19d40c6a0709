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


def train_compare(scene="zara1", quiet=False):
    """One training epoch of the reference's own loop (``ETSequencedMiniBatchTrainer.train``, utils/trainer.py:117-150:
    sequential scenes, losses accumulated over ``batch_size`` scenes, backward, gradient clipping, AdamW step) from the
    same ``state_dict``: twice with the reference's modules (the second run is the noise floor of the predictor's own GPU
    kernels) and once with ``eigentrajectory_b200`` swapped in -- fused forward, fused losses, sparse arg-min backward.
    Compares the epoch's mean training loss and the predictor's weights after the epoch."""
    from oracle import ref_loader
    import eigentrajectory_b200 as et
    ref_ET, ref_utils, ref_baseline = ref_loader.load("EigenTrajectory", "utils", "baseline")
    import utils.trainer as trainer_mod

    cfg = os.path.join(ref_loader.REF_ROOT, "config", "eigentrajectory-{baseline}-" + scene + ".json")
    hp = ref_utils.get_exp_config(cfg)
    hp.baseline = "sgcn"
    hp.dataset_dir = os.path.join(ref_loader.REF_ROOT, "datasets") + "/"
    hp.checkpoint_dir = "/tmp/et_config4_ckpt"

    ours = build_trainer(ref_utils, ref_baseline, trainer_mod, et.EigenTrajectory, hp)
    obs = torch.cat([ours.loader_train.dataset.obs_traj, ours.loader_val.dataset.obs_traj], dim=0)
    pred = torch.cat([ours.loader_train.dataset.pred_traj, ours.loader_val.dataset.pred_traj], dim=0)
    obs, pred = ref_utils.augment_trajectory(obs, pred)
    ours.model.calculate_parameters(obs.cuda(), pred.cuda())
    state = {k: v.detach().clone() for k, v in ours.model.state_dict().items()}

    def epoch(model_cls):
        trainer = build_trainer(ref_utils, ref_baseline, trainer_mod, model_cls, hp)       # fresh AdamW state
        missing = trainer.model.load_state_dict(state, strict=False)
        assert not missing.missing_keys and not missing.unexpected_keys, missing
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stderr(io.StringIO()):
            trainer.train(0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        weights = torch.cat([p.detach().flatten().double().cpu() for p in trainer.model.baseline_model.parameters()])
        return trainer.log["train_loss"][-1], weights, dt, len(trainer.loader_train)

    loss_a, w_a, t_a, scenes = epoch(ref_ET.EigenTrajectory)
    loss_a2, w_a2, _, _ = epoch(ref_ET.EigenTrajectory)
    launches = et.launch_count()
    loss_b, w_b, t_b, _ = epoch(et.EigenTrajectory)
    launches = et.launch_count() - launches
    w0 = torch.cat([v.flatten().double().cpu() for k, v in state.items() if k.startswith("baseline_model.") and v.dtype.is_floating_point])
    scale = float(w_a.abs().max())
    out = {"config": "configs[3]: one ET-SGCN training epoch, " + scene + " train split", "scenes": scenes,
           "optimizer_steps": (scenes + hp.batch_size - 1) // hp.batch_size,
           "train_loss_reference": loss_a, "train_loss_ours": loss_b, "train_loss_rel_diff": abs(loss_a - loss_b) / abs(loss_a),
           "train_loss_rel_diff_reference_rerun": abs(loss_a - loss_a2) / abs(loss_a),
           "weights_max_abs_diff": float((w_a - w_b).abs().max()), "weights_max_abs_diff_reference_rerun": float((w_a - w_a2).abs().max()),
           "weights_max_abs": scale, "weights_moved_by": float((w_a - w0).abs().max()) if w0.numel() == w_a.numel() else None,
           "ms_per_scene_reference_l2": 1e3 * t_a / scenes, "ms_per_scene_ours": 1e3 * t_b / scenes, "speedup": t_a / t_b,
           "library_launches_per_scene": launches / scenes}
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


def grad_compare(scene="zara1", n_scenes=128, quiet=False):
    """Gradient of the FIRST optimizer step of that loop (losses of the first ``batch_size`` = 128 training scenes
    accumulated exactly as utils/trainer.py:124-141 does, one backward) with respect to every SGCN weight: reference
    modules vs this package, same weights.  This is the parity statement for training -- after the optimizer steps AdamW's
    g / sqrt(v) normalisation turns last-bit gradient noise on near-zero gradients into O(lr) weight differences."""
    from oracle import ref_loader
    import eigentrajectory_b200 as et
    ref_ET, ref_utils, ref_baseline = ref_loader.load("EigenTrajectory", "utils", "baseline")
    import utils.trainer as trainer_mod

    cfg = os.path.join(ref_loader.REF_ROOT, "config", "eigentrajectory-{baseline}-" + scene + ".json")
    hp = ref_utils.get_exp_config(cfg)
    hp.baseline = "sgcn"
    hp.dataset_dir = os.path.join(ref_loader.REF_ROOT, "datasets") + "/"
    hp.checkpoint_dir = "/tmp/et_config4_ckpt"
    ours = build_trainer(ref_utils, ref_baseline, trainer_mod, et.EigenTrajectory, hp)
    obs = torch.cat([ours.loader_train.dataset.obs_traj, ours.loader_val.dataset.obs_traj], dim=0)
    pred = torch.cat([ours.loader_train.dataset.pred_traj, ours.loader_val.dataset.pred_traj], dim=0)
    obs, pred = ref_utils.augment_trajectory(obs, pred)
    ours.model.calculate_parameters(obs.cuda(), pred.cuda())
    state = {k: v.detach().clone() for k, v in ours.model.state_dict().items()}

    def first_step(trainer):
        trainer.model.load_state_dict(state, strict=False)
        trainer.model.train()
        trainer.optimizer.zero_grad()
        total, losses = None, []
        for cnt, batch in enumerate(trainer.loader_train):
            if cnt >= n_scenes:
                break
            o, p_ = [t.cuda(non_blocking=True) for t in batch[:2]]
            out = trainer.model(o, p_)
            loss = out["loss_eigentraj"] + out["loss_euclidean_ade"] + out["loss_euclidean_fde"]
            loss[torch.isnan(loss)] = 0
            losses.append(float(loss))
            total = loss if total is None else total + loss
        (total / n_scenes).backward()
        grads = torch.cat([p.grad.detach().flatten().double().cpu() for p in trainer.model.baseline_model.parameters()])
        return grads, np.asarray(losses)

    g_b, l_b = first_step(ours)
    ref = build_trainer(ref_utils, ref_baseline, trainer_mod, ref_ET.EigenTrajectory, hp)
    g_a, l_a = first_step(ref)
    g_a2, _ = first_step(ref)
    out = {"config": f"configs[3]: gradient of the first optimizer step ({n_scenes} scenes), {scene} train split",
           "parameters": int(g_a.numel()), "grad_max_abs": float(g_a.abs().max()),
           "grad_max_abs_diff": float((g_a - g_b).abs().max()), "grad_rel_max": float((g_a - g_b).abs().max() / g_a.abs().max()),
           "grad_rel_fro": float((g_a - g_b).norm() / g_a.norm()), "grad_rel_fro_reference_rerun": float((g_a - g_a2).norm() / g_a.norm()),
           "per_scene_loss_max_rel_diff": float(np.abs(l_a - l_b).max() / np.abs(l_a).max())}
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "grad":
        grad_compare(*(sys.argv[2:3]))
    elif len(sys.argv) > 1 and sys.argv[1] == "train":
        train_compare(*(sys.argv[2:3]))
    else:
        main(*(sys.argv[1:2]))
