"""CPU: the oracle against the UNMODIFIED reference staged under oracle/_ref (oracle/make_ref.sh) on fresh seeded inputs
-- beyond the frozen fixtures of tests/golden, the restatement is checked against the reference's own classes run here.
Skipped where the reference is not staged."""
import numpy as np
import pytest
import torch

from oracle import et_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (run oracle/make_ref.sh)")
HP = dict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3, obs_svd=True, pred_svd=True)


@pytest.fixture(scope="module")
def ref():
    ET, utils = ref_loader.load("EigenTrajectory", "utils")
    return ET, utils


def test_manifest_matches_staged_files():
    import hashlib
    import os
    root = ref_loader.REF_ROOT
    lines = open(os.path.join(os.path.dirname(root), "MANIFEST.sha256")).read().splitlines()
    assert len(lines) > 50
    for line in lines:
        digest, rel = line.split(None, 1)
        with open(os.path.join(root, rel.strip()), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel


def test_descriptor_round_trip_equals_reference(ref):
    ET, utils = ref
    from EigenTrajectory.descriptor import ETDescriptor
    obs, pred = O.synthetic_trajectories(20_000, seed=77)
    d = ETDescriptor(utils.DotDict(HP))
    with torch.no_grad():
        pred_norm, U_pred = d.parameter_initialization(obs, pred)
        mine = O.parameter_initialization(obs, pred, 6)
        assert torch.equal(pred_norm, mine["pred_norm"]) and torch.equal(U_pred, mine["U_pred"])
        C_obs, C_pred = d.projection(obs, pred)
        rec_obs = d.denormalize_trajectory(d.to_Euclidean_space(C_obs, d.U_obs_trunc))
        rec_pred = d.denormalize_trajectory(d.to_Euclidean_space(C_pred, d.U_pred_trunc))
        o = O.project_reconstruct(obs, pred, mine["U_obs"], mine["U_pred"])
        for a, b in zip((rec_obs, rec_pred, C_obs, C_pred), o):
            assert torch.equal(a, b)
        C20 = torch.randn(6, 20_000, 20, generator=torch.Generator().manual_seed(1))
        assert torch.equal(d.reconstruction(C20), O.descriptor_reconstruction(C20, mine["U_pred"], mine["state"]))


def test_kmeans_and_metrics_equal_reference(ref):
    ET, utils = ref
    from EigenTrajectory.kmeans import BatchKMeans
    gen = torch.Generator().manual_seed(5)
    data = (torch.randn(2, 6, 30_000, generator=gen) * torch.linspace(4, 0.3, 6)[None, :, None]).contiguous()
    km = BatchKMeans(n_clusters=20, max_iter=7)
    np.random.seed(0)
    labels = km.fit(data)
    np.random.seed(0)
    o_labels, o_cent, o_it, _ = O.kmeans_fit(data, 20, max_iter=7)
    assert torch.equal(labels, o_labels) and torch.equal(km.centroids, o_cent)
    gt = torch.randn(500, 12, 2, generator=gen).cumsum(1)
    pred = gt[None] + torch.randn(20, 500, 12, 2, generator=gen) * 0.4
    ade, fde, _ = O.ade_fde(pred, gt)
    assert np.array_equal(utils.compute_batch_ade(pred, gt), ade.numpy())
    assert np.array_equal(utils.compute_batch_fde(pred, gt), fde.numpy())
