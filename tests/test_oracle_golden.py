"""CPU: the oracle restatement vs outputs frozen from the UNMODIFIED reference (tests/golden)."""
import os

import numpy as np
import torch

from conftest import load_golden, rel_fro, rel_max, t
from oracle import et_oracle as O


def test_normaliser_and_projection_bit_exact():
    g = load_golden("descriptor_syn")
    obs, pred = t(g["obs"]), t(g["pred"])
    for tag, sca in (("sca1", True), ("sca0", False)):
        ori, rot, s = O.norm_params(obs, True, True, sca)
        assert torch.equal(ori, t(g[f"ori_{tag}"]))
        assert torch.equal(rot, t(g[f"rot_{tag}"]))
        if sca:
            assert torch.equal(s, t(g[f"sca_{tag}"]))
        assert torch.equal(O.normalize(obs, ori, rot, s), t(g[f"obs_norm_{tag}"]))
        assert torch.equal(O.normalize(pred, ori, rot, s), t(g[f"pred_norm_{tag}"]))
        C_obs, C_pred, _ = O.descriptor_projection(obs, pred, t(g[f"U_obs_{tag}"]), t(g[f"U_pred_{tag}"]), True, True, sca)
        # sgemm blocking depends on the host's thread count: matmul outputs agree to rounding, not bitwise
        assert rel_max(C_obs, g[f"C_obs_{tag}"]) < 2e-6
        assert rel_max(C_pred, g[f"C_pred_{tag}"]) < 2e-6


def test_svd_basis_matches_reference():
    g = load_golden("descriptor_syn")
    for tag in ("sca1", "sca0"):
        for side, T in (("obs", 16), ("pred", 24)):
            U, S, V = O.svd_basis(t(g[f"{side}_norm_{tag}"]), 6)
            assert rel_max(S, g[f"S_{side}_{tag}"]) < 1e-6
            assert rel_max(U, g[f"U_{side}_{tag}"]) < 1e-4          # same LAPACK call; threading may differ
            assert rel_fro(V, g[f"V_{side}_{tag}"]) < 1e-4


def test_round_trip_and_reconstruction():
    g = load_golden("descriptor_syn")
    obs, pred = t(g["obs"]), t(g["pred"])
    for tag, sca in (("sca1", True), ("sca0", False)):
        Uo, Up = t(g[f"U_obs_{tag}"]), t(g[f"U_pred_{tag}"])
        ro, rp, _, _ = O.project_reconstruct(obs, pred, Uo, Up, True, True, sca)
        assert rel_max(ro, g[f"rec_obs_{tag}"]) < 2e-6
        assert rel_max(rp, g[f"rec_pred_{tag}"]) < 2e-6
        state = O.norm_params(obs[:100], True, True, sca)
        C_in = t(g["C_in"])
        rec = O.descriptor_reconstruction(C_in, Up, state)
        assert rel_max(rec, g[f"recon20_{tag}"]) < 2e-6
        Cg = C_in.clone().requires_grad_(True)
        rec2 = O.descriptor_reconstruction(O.anchor_add(t(g[f"anchor_{tag}"]), Cg), Up, state)
        assert rel_max(rec2.detach(), g[f"recon20_anchor_{tag}"]) < 2e-6
        (rec2 * t(g[f"grad_w_{tag}"])).sum().backward()
        assert rel_max(Cg.grad, g[f"grad_C_{tag}"]) < 1e-6


def test_eth_rank_k_errors():
    g = load_golden("eth_test")
    rows, (Uo, So, Up, Sp) = O.rank_k_errors(t(g["obs"]), t(g["pred"]))
    assert np.allclose([r[1] for r in rows], g["err_obs"], rtol=1e-6, atol=0)
    assert np.allclose([r[2] for r in rows], g["err_pred"], rtol=1e-6, atol=0)
    # SURVEY section 4 table, k = 6
    assert abs(rows[5][1] - 0.0269) < 5e-5 and abs(rows[5][2] - 0.0654) < 5e-5
    assert rel_max(Sp[:6], [184.413, 26.252, 14.566, 7.138, 4.352, 2.989]) < 1e-5


def test_kmeans_trace_bit_exact():
    g = load_golden("kmeans")
    data = t(g["data"])
    c = O.kmeans_farthest_init(data, 20, int(g["first_index"]))
    assert torch.equal(c, t(g["init_centroids"]))
    for it in range(12):
        ms, lb = O.kmeans_assign(data, c)
        assert torch.equal(lb, t(g["trace_labels"][it]).long())
        assert torch.equal(ms, t(g["trace_maxsims"][it]))
        c = O.kmeans_update(data, lb, 20)
        assert torch.equal(c, t(g["trace_centroids"][it + 1]))
    labels, cent, n_it, _ = O.kmeans_fit(data, 20, first_index=int(g["first_index"]))
    assert torch.equal(labels, t(g["fit_labels"]).long())
    assert torch.equal(cent, t(g["fit_centroids"]))
    for n in (1, 31, 33, 1000):
        ms, lb = O.kmeans_assign(t(g[f"edge{n}_data"]), t(g[f"edge{n}_cent"]))
        assert torch.equal(lb, t(g[f"edge{n}_labels"]))
        assert torch.equal(ms, t(g[f"edge{n}_maxsims"]))


def test_plain_c_kmeans_assign_bit_exact():
    """oracle/et_oracle_kmeans.c (gcc, no contraction) against the reference's lock-step trace, its ragged-size
    fixtures, and the torch restatement on other (d, K, N) shapes of the summation-order rule -- all bit for bit."""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", root, "oracle"], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(root, "oracle", "_build", "libet_oracle.so"))
    vp = ctypes.c_void_p
    lib.et_oracle_kmeans_assign.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, vp, vp]
    lib.et_oracle_kmeans_assign.restype = None

    def c_assign(data, cent):
        data, cent = data.contiguous().float(), cent.contiguous().float()
        l, d, n = data.shape
        labels = torch.empty((l, n), dtype=torch.int64)
        sims = torch.empty((l, n), dtype=torch.float32)
        lib.et_oracle_kmeans_assign(data.data_ptr(), cent.data_ptr(), l, d, n, cent.size(-1), labels.data_ptr(), sims.data_ptr())
        return sims, labels

    g = load_golden("kmeans")
    data = t(g["data"])
    for it in range(12):
        ms, lb = c_assign(data, t(g["trace_centroids"][it]))
        assert torch.equal(lb, t(g["trace_labels"][it]).long()), it
        assert torch.equal(ms, t(g["trace_maxsims"][it])), it
    for n in (1, 31, 33, 1000):
        ms, lb = c_assign(t(g[f"edge{n}_data"]), t(g[f"edge{n}_cent"]))
        assert torch.equal(lb, t(g[f"edge{n}_labels"])) and torch.equal(ms, t(g[f"edge{n}_maxsims"]))
    gen = torch.Generator().manual_seed(11)
    for d, k, n in ((2, 5, 1000), (3, 7, 777), (6, 40, 5000), (8, 33, 4099), (16, 64, 3000), (5, 20, 6), (6, 4, 100), (6, 20, 70)):
        x, c = torch.randn(2, d, n, generator=gen), torch.randn(2, d, k, generator=gen)
        ms, lb = c_assign(x, c)
        o_ms, o_lb = O.kmeans_assign(x, c)
        assert torch.equal(lb, o_lb) and torch.equal(ms, o_ms), (d, k, n)
    # NaN centroid (an emptied cluster): torch.max lets the first NaN win
    x, c = torch.randn(1, 6, 50, generator=gen), torch.randn(1, 6, 20, generator=gen)
    c[0, :, 7] = float("nan")
    ms, lb = c_assign(x, c)
    o_ms, o_lb = O.kmeans_assign(x, c)
    assert torch.equal(lb, o_lb) and bool(torch.isnan(ms).all()) and bool((lb == 7).all())


def test_metrics_bit_exact():
    g = load_golden("metrics")
    ade, fde, _ = O.ade_fde(t(g["pred"]), t(g["gt"]))
    assert np.array_equal(ade.numpy(), g["ade"])
    assert np.array_equal(fde.numpy(), g["fde"])


def test_tcc_col_bit_exact():
    g = load_golden("metrics")
    assert np.array_equal(O.tcc(t(g["pred"]), t(g["gt"])).numpy(), g["tcc"])
    assert np.array_equal(O.col(t(g["pred"])).numpy(), g["col"])
    assert np.array_equal(O.col(t(g["scene_pred"])).numpy(), g["scene_col"])
    assert np.array_equal(O.tcc(t(g["scene_pred"]), t(g["scene_gt"])[None]).numpy(), g["scene_tcc"])
    assert 0 < g["scene_col"].mean() < 100          # the crowded fixture really has collisions


def test_eth_init_spectra_table():
    """frozen singular values of the ETH init matrices; fp32 LAPACK vs fp64 noise floor."""
    g = load_golden("eth_init")
    assert int(g["n_total"]) == 70316 and int(g["n_moving"]) == 14456
    assert rel_max(g["S_obs_m"], [1393.2605, 133.5286, 39.0587, 19.2903, 11.2614, 6.4631]) < 1e-6
    assert rel_max(g["S_pred_m"], [2827.5923, 569.8727, 122.9585, 102.1002, 37.7803, 35.3952]) < 1e-6
    for tag in ("m", "s"):
        for side in ("obs", "pred"):
            assert rel_max(g[f"S_{side}_{tag}"], g[f"S_{side}64_{tag}"]) < 1e-5


def test_kmeans_update_empty_cluster_is_nan():
    data = torch.randn(1, 6, 50)
    labels = torch.zeros(1, 50, dtype=torch.long)
    c = O.kmeans_update(data, labels, 3)
    assert torch.isnan(c[0, :, 1:]).all() and not torch.isnan(c[0, :, 0]).any()
