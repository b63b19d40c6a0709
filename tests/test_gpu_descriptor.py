"""GPU parity: normaliser / projection / reconstruction / fused round trip vs the frozen reference
outputs (tests/golden) and vs the CPU oracle on seeded inputs.  Everything goes through the C ABI.

Tolerance (BASELINE.json north_star): 1e-5 relative fp32, stated as max|x - ref| / max|ref| and as
relative Frobenius norm.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_fro, rel_max, t

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def et():
    import eigentrajectory_b200 as et
    et.load_library()
    return et


@pytest.fixture(scope="module")
def O():
    from oracle import et_oracle
    return et_oracle


HP = dict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.419, obs_svd=True, pred_svd=True)


def make_desc(et, g, tag, sca, device="cuda"):
    d = et.ETDescriptor(et.DotDict(HP), norm_sca=sca).to(device)
    d.U_obs_trunc.data = t(g[f"U_obs_{tag}"]).to(device)
    d.U_pred_trunc.data = t(g[f"U_pred_{tag}"]).to(device)
    return d


@pytest.mark.parametrize("tag,sca", [("sca1", True), ("sca0", False)])
def test_trajnorm_matches_reference(et, tag, sca):
    g = load_golden("descriptor_syn")
    obs, pred = t(g["obs"]).cuda(), t(g["pred"]).cuda()
    tn = et.TrajNorm(ori=True, rot=True, sca=sca)
    tn.calculate_params(obs)
    assert torch.equal(tn.traj_ori.cpu(), t(g[f"ori_{tag}"]))
    assert (tn.traj_rot.cpu() - t(g[f"rot_{tag}"])).abs().max() < 5e-7
    if sca:
        assert rel_max(tn.traj_sca.cpu(), g[f"sca_{tag}"]) < 1e-6
    on, pn = tn.normalize(obs), tn.normalize(pred)
    assert rel_max(on.cpu(), g[f"obs_norm_{tag}"]) < TOL
    assert rel_max(pn.cpu(), g[f"pred_norm_{tag}"]) < TOL
    back = tn.denormalize(pn)
    assert rel_max(back.cpu(), g["pred"]) < TOL
    # host tensors in -> host tensors out, same numbers
    tn2 = et.TrajNorm(ori=True, rot=True, sca=sca)
    tn2.calculate_params(t(g["obs"]))
    assert not tn2.traj_ori.is_cuda and torch.equal(tn2.normalize(t(g["pred"])), pn.cpu())


@pytest.mark.parametrize("tag,sca", [("sca1", True), ("sca0", False)])
def test_projection_matches_reference(et, tag, sca):
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, tag, sca)
    obs, pred = t(g["obs"]).cuda(), t(g["pred"]).cuda()
    C_obs, C_pred = d.projection(obs, pred)
    assert C_obs.shape == (6, 1536) and C_pred.shape == (6, 1536) and not C_obs.requires_grad
    assert rel_max(C_obs.cpu(), g[f"C_obs_{tag}"]) < TOL and rel_fro(C_obs.cpu(), g[f"C_obs_{tag}"]) < TOL
    assert rel_max(C_pred.cpu(), g[f"C_pred_{tag}"]) < TOL and rel_fro(C_pred.cpu(), g[f"C_pred_{tag}"]) < TOL
    tn = d.traj_normalizer
    assert tn.traj_ori.shape == (1536, 1, 2) and tn.traj_rot.shape == (1536, 2, 2)
    assert torch.equal(tn.traj_ori.cpu(), t(g[f"ori_{tag}"]))
    C_only, none = d.projection(obs)
    assert none is None and torch.equal(C_only, C_obs)
    # to_ET_space / to_Euclidean_space on already normalised data (generic kernels)
    C2 = d.to_ET_space(t(g[f"pred_norm_{tag}"]).cuda(), d.U_pred_trunc)
    assert rel_max(C2.cpu(), g[f"C_pred_{tag}"]) < TOL
    rec_norm = d.to_Euclidean_space(C2, d.U_pred_trunc)
    rec = d.denormalize_trajectory(rec_norm)
    assert rel_max(rec.cpu(), g[f"rec_pred_{tag}"]) < TOL


@pytest.mark.parametrize("n", [4096, 4097, 8192, 100_000, 100_002])
@pytest.mark.parametrize("sca", [True, False])
def test_projection_large_batches_vs_oracle(et, O, n, sca):
    """n >= 4096 with n % 4 == 0 takes the TMA pipeline (coefficients by tensor stores, state by coalesced stores);
    other sizes the thread-per-row kernel.  Both against the oracle, including the stored normaliser state."""
    g = load_golden("descriptor_syn")
    tag = "sca1" if sca else "sca0"
    d = make_desc(et, g, tag, sca)
    obs, pred = O.synthetic_trajectories(n, seed=n)
    C_obs, C_pred = d.projection(obs.cuda(), pred.cuda())
    o_co, o_cp, (o_ori, o_rot, o_sca) = O.descriptor_projection(obs, pred, t(g[f"U_obs_{tag}"]), t(g[f"U_pred_{tag}"]), True, True, sca)
    assert rel_max(C_obs.cpu(), o_co) < TOL and rel_max(C_pred.cpu(), o_cp) < TOL
    tn = d.traj_normalizer
    assert torch.equal(tn.traj_ori.cpu(), o_ori) and (tn.traj_rot.cpu() - o_rot).abs().max() < 5e-7
    if sca:
        assert rel_max(tn.traj_sca.cpu(), o_sca) < 1e-6
    # the state written by the fused kernel drives reconstruction of the same batch
    C20 = torch.randn(6, n, 20, generator=torch.Generator().manual_seed(1)) * 0.3
    if n <= 8192:
        rec = d.reconstruction(C20.cuda())
        o_rec = O.descriptor_reconstruction(C20, t(g[f"U_pred_{tag}"]), (o_ori, o_rot, o_sca))
        assert rel_max(rec.cpu(), o_rec) < TOL


@pytest.mark.parametrize("tag,sca", [("sca1", True), ("sca0", False)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_project_reconstruct_matches_reference(et, tag, sca, variant):
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, tag, sca)
    obs, pred = t(g["obs"]).cuda(), t(g["pred"]).cuda()
    ro, rp, co, cp = d.project_reconstruct(obs, pred, variant=variant)
    for mine, key in ((ro, "rec_obs"), (rp, "rec_pred"), (co, "C_obs"), (cp, "C_pred")):
        ref = g[f"{key}_{tag}"]
        assert rel_max(mine.cpu(), ref) < TOL, (key, variant)
        assert rel_fro(mine.cpu(), ref) < TOL, (key, variant)
    ro2, rp2, none_o, none_p = d.project_reconstruct(obs, pred, want_coeffs=False, variant=variant)
    assert none_o is None and none_p is None
    assert torch.equal(ro2, ro) and torch.equal(rp2, rp)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 128, 129, 1000, 4097, 50000])
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_project_reconstruct_ragged_sizes_vs_oracle(et, O, n, variant):
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, "sca1", True)
    obs, pred = O.synthetic_trajectories(n, seed=100 + n)
    ro, rp, co, cp = d.project_reconstruct(obs.cuda(), pred.cuda(), variant=variant)
    o_ro, o_rp, o_co, o_cp = O.project_reconstruct(obs, pred, t(g["U_obs_sca1"]), t(g["U_pred_sca1"]))
    for mine, ref in ((ro, o_ro), (rp, o_rp), (co, o_co), (cp, o_cp)):
        assert mine.shape == ref.shape
        assert rel_max(mine.cpu(), ref) < TOL, (n, variant)


def test_project_reconstruct_empty_and_errors(et):
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, "sca1", True)
    ro, rp, co, cp = d.project_reconstruct(torch.zeros(0, 8, 2).cuda(), torch.zeros(0, 12, 2).cuda())
    assert ro.shape == (0, 8, 2) and cp.shape == (6, 0)
    with pytest.raises(et.ETLibraryError):
        et.ops.project_reconstruct(torch.zeros(4, 2, 2).cuda(), torch.zeros(4, 12, 2).cuda(), torch.zeros(4, 2).cuda(),
                                   torch.zeros(24, 2).cuda())        # T_obs < 3


def test_project_reconstruct_generic_shapes_vs_oracle(et, O):
    """Shapes off the (8,12,6) fast path run the generic kernels."""
    for (to, tp, k) in ((5, 7, 3), (8, 12, 4), (10, 20, 8)):
        obs, pred = O.synthetic_trajectories(777, seed=to, t_obs=to, t_pred=tp)
        ref = O.parameter_initialization(obs, pred, k)
        ro, rp, co, cp = et.ops.project_reconstruct(obs.cuda(), pred.cuda(), ref["U_obs"], ref["U_pred"])
        o = O.project_reconstruct(obs, pred, ref["U_obs"], ref["U_pred"])
        for mine, want in zip((ro, rp, co, cp), o):
            assert rel_max(mine.cpu(), want) < TOL, (to, tp, k)
        C_obs, C_pred, st = et.ops.project(obs.cuda(), pred.cuda(), ref["U_obs"], ref["U_pred"])
        assert rel_max(C_pred.cpu(), o[3]) < TOL


def test_round_trip_is_a_projection(et, O):
    """Size-independent property at the headline size: P(P(x)) = P(x) and the residual is orthogonal to U."""
    n = 1_000_000
    obs, pred = O.synthetic_trajectories(n, seed=0)
    obs, pred = obs.cuda(), pred.cuda()
    d = et.ETDescriptor(et.DotDict(HP), norm_sca=False).cuda()
    d.parameter_initialization(obs, pred)
    ro, rp, co, cp = d.project_reconstruct(obs, pred)
    ro2, rp2, co2, cp2 = d.project_reconstruct(ro, rp)
    # same normaliser state requires the same last/third-last observed frames: compare coefficients of
    # the first pass with the projection of its own reconstruction under the ORIGINAL state
    tn = et.TrajNorm(True, True, False)
    tn.calculate_params(obs)
    cp_again = d.to_ET_space(tn.normalize(rp), d.U_pred_trunc)
    assert rel_max(cp_again, cp) < 2e-5
    resid = tn.normalize(pred) - tn.normalize(rp)
    assert d.to_ET_space(resid, d.U_pred_trunc).abs().max() < 2e-4 * cp.abs().max()
    assert torch.isfinite(ro).all() and torch.isfinite(rp).all()


def test_headline_size_vs_oracle(et, O):
    """Config 2, the exact bench inputs (synthetic_trajectories(1e6, seed=0), k = 6, ori+rot+sca, basis from the same
    data): fused round trip and stand-alone projection against the CPU oracle at N = 1e6.
    Tolerance 1e-5 (north_star) as max|x - ref| / max|ref| and relative Frobenius norm."""
    n = 1_000_000
    obs, pred = O.synthetic_trajectories(n, seed=0)
    d = et.ETDescriptor(et.DotDict(HP)).cuda()
    d.parameter_initialization(obs.cuda(), pred.cuda())
    Uo, Up = d.U_obs_trunc.detach().cpu(), d.U_pred_trunc.detach().cpu()
    want = O.project_reconstruct(obs, pred, Uo, Up)
    worst = 0.0
    for variant in (0, 1):
        got = d.project_reconstruct(obs.cuda(), pred.cuda(), variant=variant)
        for name, mine, ref in zip(("rec_obs", "rec_pred", "C_obs", "C_pred"), got, want):
            e_max, e_fro = rel_max(mine.cpu(), ref), rel_fro(mine.cpu(), ref)
            worst = max(worst, e_max, e_fro)
            assert e_max < TOL and e_fro < TOL, (name, variant, e_max, e_fro)
    C_obs, C_pred = d.projection(obs.cuda(), pred.cuda())
    for name, mine, ref in (("C_obs", C_obs, want[2]), ("C_pred", C_pred, want[3])):
        e_max, e_fro = rel_max(mine.cpu(), ref), rel_fro(mine.cpu(), ref)
        worst = max(worst, e_max, e_fro)
        assert e_max < TOL and e_fro < TOL, (name, "projection", e_max, e_fro)
    # the host-buffer (pipelined H2D / kernel / D2H) path returns the same numbers as the resident one
    host = d.project_reconstruct(obs.pin_memory(), pred.pin_memory())
    resident = d.project_reconstruct(obs.cuda(), pred.cuda())
    for mine, dev in zip(host, resident):
        assert not mine.is_cuda and torch.equal(mine, dev.cpu())
    print(f"headline size vs oracle: worst relative error {worst:.3e} (tolerance {TOL:g})")


@pytest.mark.parametrize("tag,sca", [("sca1", True), ("sca0", False)])
def test_reconstruction_and_gradient_match_reference(et, tag, sca):
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, tag, sca)
    obs, pred = t(g["obs"]).cuda(), t(g["pred"]).cuda()
    d.projection(obs[:100], pred[:100])
    C_in = t(g["C_in"]).cuda()
    rec = d.reconstruction(C_in)
    assert rec.shape == (20, 100, 12, 2)
    assert rel_max(rec.cpu(), g[f"recon20_{tag}"]) < TOL and rel_fro(rec.cpu(), g[f"recon20_{tag}"]) < TOL
    assert torch.equal(d(C_in), rec)                                   # forward is an alias
    # anchor add (torch broadcast, as ETAnchor.forward) then reconstruction, with autograd
    a = et.ETAnchor(et.DotDict(HP)).cuda()
    a.C_anchor.data = t(g[f"anchor_{tag}"]).cuda()
    Cg = C_in.clone().requires_grad_(True)
    rec2 = d.reconstruction(a(Cg))
    assert rel_max(rec2.detach().cpu(), g[f"recon20_anchor_{tag}"]) < TOL
    (rec2 * t(g[f"grad_w_{tag}"]).cuda()).sum().backward()
    assert rel_max(Cg.grad.cpu(), g[f"grad_C_{tag}"]) < TOL and rel_fro(Cg.grad.cpu(), g[f"grad_C_{tag}"]) < TOL
    assert a.C_anchor.grad is None and d.U_pred_trunc.grad is None
    # fused anchor path gives the same numbers
    rec3 = d.reconstruction(C_in, anchor=a.C_anchor)
    assert rel_max(rec3.cpu(), g[f"recon20_anchor_{tag}"]) < TOL


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 100, 321, 5000])
def test_reconstruction_ragged_vs_oracle(et, O, n):
    """Fast (bulk-copy) path for n >= 32, generic kernel below; gradient by both."""
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, "sca1", True)
    obs, pred = O.synthetic_trajectories(n, seed=7 + n)
    d.projection(obs.cuda(), pred.cuda())
    gen = torch.Generator().manual_seed(n)
    C = torch.randn(6, n, 20, generator=gen) * torch.tensor([8., 3., 1., .5, .3, .2])[:, None, None]
    w = torch.randn(20, n, 12, 2, generator=gen)
    Cd = C.cuda().requires_grad_(True)
    rec = d.reconstruction(Cd)
    (rec * w.cuda()).sum().backward()
    state = O.norm_params(obs)
    Co = C.clone().requires_grad_(True)
    o_rec = O.descriptor_reconstruction(Co, t(g["U_pred_sca1"]), state)
    (o_rec * w).sum().backward()
    assert rel_max(rec.detach().cpu(), o_rec.detach()) < TOL, n
    assert rel_max(Cd.grad.cpu(), Co.grad) < TOL, n


def test_reconstruction_other_sample_counts(et, O):
    g = load_golden("descriptor_syn")
    for s in (1, 3, 20):
        hp = dict(HP, num_samples=s)
        d = et.ETDescriptor(et.DotDict(hp), norm_sca=True).cuda()
        d.U_obs_trunc.data, d.U_pred_trunc.data = t(g["U_obs_sca1"]).cuda(), t(g["U_pred_sca1"]).cuda()
        obs, pred = O.synthetic_trajectories(257, seed=s)
        d.projection(obs.cuda(), pred.cuda())
        C = torch.randn(6, 257, s)
        rec = d.reconstruction(C.cuda())
        o_rec = O.descriptor_reconstruction(C, t(g["U_pred_sca1"]), O.norm_params(obs))
        assert rel_max(rec.cpu(), o_rec) < TOL


def test_static_pedestrian_semantics(et, O):
    """|d| = 0: rotation is the identity; with sca the reference yields inf/NaN and so do we; without sca finite."""
    obs, pred = O.synthetic_trajectories(64, seed=5)
    obs[3, -3:] = obs[3, -1]                      # pedestrian 3 did not move over the last three frames
    g = load_golden("descriptor_syn")
    d0 = make_desc(et, g, "sca0", False)
    C_obs, C_pred = d0.projection(obs.cuda(), pred.cuda())
    rot = d0.traj_normalizer.traj_rot[3].cpu()
    assert torch.equal(rot, torch.eye(2))
    o_co, o_cp, _ = O.descriptor_projection(obs, pred, t(g["U_obs_sca0"]), t(g["U_pred_sca0"]), True, True, False)
    assert rel_max(C_pred.cpu(), o_cp) < TOL and torch.isfinite(C_obs).all()
    d1 = make_desc(et, g, "sca1", True)
    C1, _ = d1.projection(obs.cuda(), pred.cuda())
    assert torch.isinf(d1.traj_normalizer.traj_sca[3]).all()
    assert not torch.isfinite(C1[:, 3]).all() and torch.isfinite(C1[:, :3]).all()


def test_host_buffers_round_trip(et, O):
    """The API a reference user calls, with CPU tensors: results come back on the CPU."""
    g = load_golden("descriptor_syn")
    d = make_desc(et, g, "sca1", True, device="cpu")
    C_obs, C_pred = d.projection(t(g["obs"]), t(g["pred"]))
    assert not C_obs.is_cuda and rel_max(C_pred, g["C_pred_sca1"]) < TOL
    d.projection(t(g["obs"])[:100], t(g["pred"])[:100])
    rec = d.reconstruction(t(g["C_in"]))
    assert not rec.is_cuda and rel_max(rec, g["recon20_sca1"]) < TOL


def test_descriptor_maps_are_differentiable(et, O):
    """normalize / to_ET_space / to_Euclidean_space / denormalize carry autograd with respect to their data argument,
    as the reference's tensor algebra does (normalizer.py:42-62, descriptor.py:59-89); the state and ``evec`` are
    constants.  Gradients against torch autograd through the oracle's restatement."""
    g = load_golden("descriptor_syn")
    obs, pred = t(g["obs"])[:777], t(g["pred"])[:777]
    U = t(g["U_pred_sca1"])
    st = O.norm_params(obs)
    w1 = torch.randn(6, 777, generator=torch.Generator().manual_seed(5))
    w2 = torch.randn(777, 12, 2, generator=torch.Generator().manual_seed(6))

    def chain(x, norm, to_et, to_eu, denorm):
        C = to_et(norm(x))
        back = denorm(to_eu(C))
        return (C * w1.to(C.device)).sum() + (back * w2.to(back.device)).sum(), C, back

    x_ref = pred.clone().requires_grad_(True)
    loss_ref, C_ref, back_ref = chain(x_ref, lambda x: O.normalize(x, *st), lambda x: O.project(x, U),
                                      lambda C: O.unproject(C, U), lambda x: O.denormalize(x, *st))
    loss_ref.backward()

    tn = et.TrajNorm()
    tn.calculate_params(obs.cuda())
    d = make_desc(et, g, "sca1", True)
    x = pred.clone().cuda().requires_grad_(True)
    loss, C, back = chain(x, tn.normalize, lambda v: d.to_ET_space(v, d.U_pred_trunc),
                          lambda c: d.to_Euclidean_space(c, d.U_pred_trunc), tn.denormalize)
    assert C.requires_grad and back.requires_grad
    loss.backward()
    assert rel_max(C.detach().cpu(), C_ref.detach()) < TOL and rel_max(back.detach().cpu(), back_ref.detach()) < TOL
    assert rel_max(x.grad.cpu(), x_ref.grad) < TOL and rel_fro(x.grad.cpu(), x_ref.grad) < TOL
    assert d.U_pred_trunc.grad is None                      # evec is detached, as in the reference
    # without requires_grad (or under no_grad) nothing is recorded
    with torch.no_grad():
        assert not tn.normalize(x).requires_grad
    assert not tn.normalize(pred.cuda()).requires_grad


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device(et, O):
    """A module / tensors on cuda:1 while cuda:0 is the current device: every op runs on the tensors' device (grids,
    attributes and cooperative launches included) and the caller's current device is left untouched."""
    g = load_golden("descriptor_syn")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 1)
    obs, pred = O.synthetic_trajectories(20_000, seed=9)
    d = et.ETDescriptor(et.DotDict(HP)).to(dev)
    d.parameter_initialization(obs.to(dev), pred.to(dev))          # cooperative Gram pass + eigen-solve on cuda:1
    assert torch.cuda.current_device() == 0
    got = d.project_reconstruct(obs.to(dev), pred.to(dev))
    want = O.project_reconstruct(obs, pred, d.U_obs_trunc.detach().cpu(), d.U_pred_trunc.detach().cpu())
    for mine, ref in zip(got, want):
        assert mine.device == dev and rel_max(mine.cpu(), ref) < TOL
    km = et.BatchKMeans(n_clusters=20, max_iter=5)
    data = got[3].unsqueeze(0).contiguous()
    labels = km.fit(data, centroids=data[:, :, :20].contiguous())   # persistent cooperative kernel on cuda:1
    o_lab, _, _, _ = O.kmeans_fit(data.cpu(), 20, centroids=data[:, :, :20].cpu().contiguous(), max_iter=5)
    assert labels.device == dev and int((labels.cpu() != o_lab).sum()) <= 2
    assert torch.cuda.current_device() == 0
