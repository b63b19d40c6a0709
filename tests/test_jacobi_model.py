"""CPU: the numpy model of the eigen-solve kernel's arithmetic (oracle.jacobi_eig_model <-> eig_jacobi_fast2 in
eigentrajectory_b200/csrc/et_svd.cu) against the golden SVDs of the reference on the ETH initialisation set and against
the same solve with exactly computed rotation angles.  What it pins: the shortened rotation chain (two ~21-bit rsqrt
seeds, one Newton step, series normalisation) is orthogonal to rounding, needs the same sweeps, and lands on the
reference's basis."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import et_oracle as O      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def gram_of_group(tag):
    d = np.load(os.path.join(GOLD, "eth_init_data.npz"))
    obs, pred = torch.from_numpy(d["obs"]), torch.from_numpy(d["pred"])
    flip = torch.tensor([[[1.0, -1.0]]])
    obs, pred = torch.cat([obs, obs * flip]), torch.cat([pred, pred * flip])        # the init set is flip-augmented (utils/utils.py:79-81)
    rows = O.static_mask(obs, 0.419)                                          # static_dist of the ETH configuration (tests/golden/make_golden.py HP)
    rows = rows if tag == "m" else ~rows
    obs, pred = obs[rows], pred[rows]
    state = O.norm_params(obs, True, True, tag == "m")          # descriptor.py:125-131 (scale only for the moving group)
    out = []
    for x in (obs, pred):
        M = O.normalize(x, *state).reshape(x.shape[0], -1).double().numpy()   # fp32 normalisation, fp64 products: as et_gram
        out.append(M.T @ M)
    return out


def projector(U):
    return U @ U.T


@pytest.mark.parametrize("tag", ["m", "s"])
def test_short_chain_matches_exact_angles_and_the_reference_basis(tag):
    g = np.load(os.path.join(GOLD, "eth_init.npz"))
    for name, G in zip(("obs", "pred"), gram_of_group(tag)):
        U, S, counts, worst = O.jacobi_eig_model(G, 6)
        Ue, Se, counts_e, _ = O.jacobi_eig_model(G, 6, rotation=O.jacobi_rotation_exact)
        assert worst <= 1e-15, worst                                  # c^2 + s^2 = 1 to rounding
        assert len(counts) == len(counts_e) and counts[-1] == 0, (counts, counts_e)
        assert abs(sum(counts) - sum(counts_e)) <= 0.02 * sum(counts_e), (counts, counts_e)
        assert np.abs(U.T @ U - np.eye(6)).max() <= 1e-13
        assert np.linalg.norm(projector(U) - projector(Ue)) <= 1e-9
        assert np.abs(S - Se).max() / Se[0] <= 1e-12
        lam = S * S
        assert np.abs(G @ U - U * lam).max() / lam[0] <= 1e-13      # eigen-pairs of G itself
        # ... and of the reference: its float64 SVD of the same group (frozen by tests/golden/make_golden.py)
        U64, S64 = g[f"U_{name}64_{tag}"].astype(np.float64), g[f"S_{name}64_{tag}"].astype(np.float64)
        assert np.abs(S - S64).max() / S64[0] <= 1e-5, (name, tag, S, S64)
        floor = np.linalg.norm(projector(g[f"U_{name}_{tag}"].astype(np.float64)) - projector(U64))   # the reference's own fp32 distance
        assert np.linalg.norm(projector(U) - projector(U64)) <= max(1e-5, 2 * floor), (name, tag)


@pytest.mark.parametrize("seed_error", [2.0 ** -21, -2.0 ** -21, 2.0 ** -18])
def test_rotation_is_orthogonal_whatever_the_seed_accuracy(seed_error):
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(2000):
        app, aqq = rng.lognormal(0, 6, 2)
        apq = rng.normal() * np.sqrt(app * aqq) * 10.0 ** rng.uniform(-12, 0)
        c, s = O.jacobi_rotation_short_chain(app, aqq, apq, seed_error)
        ce, se = O.jacobi_rotation_exact(app, aqq, apq)
        worst = max(worst, abs(c * c + s * s - 1.0))
        # the angle: set by 1/h after ONE Newton step, i.e. to ~(3/2) seed_error^2-level accuracy
        assert abs(c * se - s * ce) <= 4.0 * (abs(seed_error) + 2.0 ** -21) ** 2 + 1e-15, (app, aqq, apq)
        assert abs(s) <= c * (1 + 1e-12) and c > 0                    # inner rotation: |theta| <= pi / 4
    assert worst <= 1e-15, worst


def test_rank_deficient_matrix_converges():
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.normal(size=(24, 24)))
    lam = np.logspace(0, -3, 24)
    lam[5:] = 0.0
    G = (q * lam) @ q.T
    U, S, counts, _ = O.jacobi_eig_model(G, 5)
    assert len(counts) <= 14 and counts[-1] == 0, counts
    assert np.abs(S - np.sqrt(lam[:5])).max() <= 1e-8
    assert np.linalg.norm(projector(U) - projector(q[:, :5])) <= 1e-7
