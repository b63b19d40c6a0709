"""GPU: the evaluation loop of the reference's ETTrainer.test (utils/trainer.py:172-195) end to end through this package
-- native dataset preprocessing -> loader (batch_size=1, tensors resident in HBM) -> EigenTrajectory.forward through the
predictor hook seam -> ADE / FDE / TCC / COL -- against the reference's own run frozen in tests/golden/eth_eval.npz."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-5
HP = dict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.419, obs_svd=True, pred_svd=True)


@pytest.fixture(scope="module")
def et():
    import eigentrajectory_b200 as et
    et.load_library()
    return et


def build_model(et, mf):
    W = torch.from_numpy(mf["W"])

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.W = torch.nn.Parameter(W.clone())

        def forward(self, x):                     # x (8,N) -> (6,N,20)
            return (self.W @ x).reshape(6, 20, -1).permute(0, 2, 1)

    hook = types.SimpleNamespace(
        model_forward_pre_hook=lambda C, o, info=None: torch.cat([C, o], dim=0),
        model_forward=lambda x, m: m(x),
        model_forward_post_hook=lambda y, info=None: y)
    model = et.EigenTrajectory(Stub(), hook, et.DotDict(dict(HP))).cuda()
    sd = {k[3:]: torch.from_numpy(mf[k]) for k in mf.files if k.startswith("sd_")}
    assert not model.load_state_dict(sd, strict=False).unexpected_keys
    return model.eval()


@pytest.mark.parametrize("resident", [True, False])
def test_eth_test_loop_matches_reference(et, tmp_path, resident):
    from eigentrajectory_b200 import dataloader as dl
    gold, ds, mf = load_golden("eth_eval"), load_golden("dataset"), load_golden("model_forward")
    d = tmp_path / "eth" / "test"
    d.mkdir(parents=True)
    (d / "biwi_eth.txt").write_bytes(ds["eth_text_biwi_eth.txt"].tobytes())
    loader = dl.get_dataloader(str(tmp_path / "eth"), "test", 8, 12, batch_size=1, device="cuda" if resident else None)
    model = build_model(et, mf)
    funcs = {"ADE": et.compute_batch_ade, "FDE": et.compute_batch_fde, "TCC": et.compute_batch_tcc, "COL": et.compute_batch_col}
    vals = {k: [] for k in funcs}
    launches = et.launch_count()
    scenes = 0
    with torch.no_grad():
        for batch in loader:
            obs, pred = [x.cuda(non_blocking=True) for x in batch[:2]]
            assert batch[0].is_cuda == resident
            out = model(obs)
            before = et.launch_count()
            for k, fn in funcs.items():
                v = fn(out["recon_traj"], pred)
                assert isinstance(v, np.ndarray) and v.shape == (obs.size(0),)
                vals[k].append(v)
            # ADE, FDE and TCC called in sequence (utils/trainer.py:186-193) share ONE pass over the samples; COL is the
            # second launch
            assert et.launch_count() - before == 2, et.launch_count() - before
            scenes += 1
    assert et.launch_count() - launches == 4 * scenes        # forward: project + reconstruct; metrics: one pass + COL
    for k in funcs:
        mine, ref = np.concatenate(vals[k]), gold[f"per_ped_{k}"]
        assert mine.shape == ref.shape == (181,)
        if k == "COL":                                     # percentages of colliding samples: exact counts
            assert np.array_equal(mine, ref), k
        else:
            assert np.abs(mine - ref).max() <= TOL * max(1.0, float(np.abs(ref).max())), (k, np.abs(mine - ref).max())
        assert abs(float(mine.mean()) - float(gold[f"mean_{k}"])) <= TOL * max(1.0, abs(float(gold[f"mean_{k}"]))), k
    # one fused pass gives the same four metrics
    obs, pred = [x.cuda() for x in next(iter(loader))[:2]]
    with torch.no_grad():
        rec = model(obs)["recon_traj"]
    ade, fde, col, tcc = et.compute_batch_metric(rec, pred)          # the reference's order (metrics.py:30-70)
    assert np.array_equal(ade.cpu().numpy(), et.compute_batch_ade(rec, pred))
    assert np.array_equal(col.cpu().numpy(), et.compute_batch_col(rec, pred))
