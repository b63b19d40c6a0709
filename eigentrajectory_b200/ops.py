"""Functional layer over the C ABI: torch tensors in, torch tensors out, every FLOP in libet_b200.so.

All functions accept CUDA or host tensors.  Host tensors are copied to the current CUDA device,
processed there and the results copied back (this is the end-to-end path ``bench.py`` reports as
``e2e``); nothing here computes on the CPU.  PyTorch is used for memory, streams and autograd
plumbing only.
"""
from __future__ import annotations

from ctypes import c_void_p as C_void_p

import torch

from . import _lib
from ._lib import ET_NORM_ORI, ET_NORM_ROT, ET_NORM_SCA, check, load, ptr, stream_of


# ----------------------------------------------------------------------------------------
# device plumbing
# ----------------------------------------------------------------------------------------
def compute_device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(t, dtype=torch.float32):
    """Contiguous tensor of ``dtype`` on the compute device (no copy if it already is one)."""
    if t is None:
        return None
    if t.is_cuda:
        dev = t.device
    else:
        dev = compute_device()
    return t.detach().to(device=dev, dtype=dtype, non_blocking=True).contiguous()


def back_to(t, like):
    """Return ``t`` on the device ``like`` lives on."""
    if t is None or like is None or t.device == like.device:
        return t
    return t.to(like.device)


def norm_flags(ori, rot, sca):
    return (ET_NORM_ORI if ori else 0) | (ET_NORM_ROT if rot else 0) | (ET_NORM_SCA if sca else 0)


def _ntc(traj):
    assert traj.dim() == 3 and traj.size(2) == 2, f"expected (N, T, 2), got {tuple(traj.shape)}"
    return traj.size(0), traj.size(1)


# ----------------------------------------------------------------------------------------
# normaliser (EigenTrajectory/normalizer.py)
# ----------------------------------------------------------------------------------------
def norm_params(obs, ori=True, rot=True, sca=True):
    """TrajNorm.calculate_params: returns (ori (N,1,2)|None, rot (N,2,2)|None, sca (N,1,1)|None)."""
    x = to_dev(obs)
    n, t = _ntc(x)
    o = torch.empty((n, 1, 2), device=x.device) if ori else None
    r = torch.empty((n, 2, 2), device=x.device) if rot else None
    s = torch.empty((n, 1, 1), device=x.device) if sca else None
    check(load().et_norm_params(ptr(x), n, t, norm_flags(ori, rot, sca), ptr(o), ptr(r), ptr(s), stream_of(x.device)),
          "et_norm_params")
    return back_to(o, obs), back_to(r, obs), back_to(s, obs)


def _apply_norm(fn_name, traj, ori, rot, sca):
    x = to_dev(traj)
    n, t = _ntc(x)
    o, r, s = (to_dev(v).to(x.device) if v is not None else None for v in (ori, rot, sca))
    for name, v, shape in (("ori", o, (n, 1, 2)), ("rot", r, (n, 2, 2)), ("sca", s, (n, 1, 1))):
        assert v is None or tuple(v.shape) == shape, f"normaliser state {name} has shape {tuple(v.shape)}, expected {shape}"
    out = torch.empty_like(x)
    fn = getattr(load(), fn_name)
    check(fn(ptr(x), n, t, norm_flags(o is not None, r is not None, s is not None), ptr(o), ptr(r), ptr(s), ptr(out),
             stream_of(x.device)), fn_name)
    return back_to(out, traj)


def _wants_grad(t):
    return t is not None and t.requires_grad and torch.is_grad_enabled()


class _AffineMap(torch.autograd.Function):
    """``normalize`` / ``denormalize`` with autograd with respect to the trajectory (the reference's tensor algebra is
    differentiable there, normalizer.py:42-62).  Both maps are affine in ``traj``; the vector-Jacobian product of one is
    the linear part of the other with the scale inverted: d normalize = (g * sca) @ R^T, d denormalize = (g @ R) / sca.
    The normaliser state is treated as a constant (it is derived from the observed frames once per batch)."""

    @staticmethod
    def forward(ctx, traj, fn_name, ori, rot, sca):
        ctx.fn_name = fn_name
        ctx.save_for_backward(*(v for v in (rot, sca) if v is not None))
        ctx.have = (rot is not None, sca is not None)
        return _apply_norm(fn_name, traj, ori, rot, sca)

    @staticmethod
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        rot = saved.pop(0) if ctx.have[0] else None
        sca = saved.pop(0) if ctx.have[1] else None
        inv = (1.0 / sca) if sca is not None else None
        back = "et_denormalize" if ctx.fn_name == "et_normalize" else "et_normalize"
        if rot is None and inv is None:
            return g, None, None, None, None
        return _apply_norm(back, g.contiguous(), None, rot, inv), None, None, None, None


def normalize(traj, ori, rot, sca):
    """TrajNorm.normalize with explicit state (None = stage disabled); differentiable with respect to ``traj``."""
    if _wants_grad(traj):
        return _AffineMap.apply(traj, "et_normalize", ori, rot, sca)
    return _apply_norm("et_normalize", traj, ori, rot, sca)


def denormalize(traj, ori, rot, sca):
    """TrajNorm.denormalize with explicit state (None = stage disabled); differentiable with respect to ``traj``."""
    if _wants_grad(traj):
        return _AffineMap.apply(traj, "et_denormalize", ori, rot, sca)
    return _apply_norm("et_denormalize", traj, ori, rot, sca)


# ----------------------------------------------------------------------------------------
# descriptor (EigenTrajectory/descriptor.py)
# ----------------------------------------------------------------------------------------
def _to_et_space_raw(traj, evec):
    U = to_dev(evec)
    x = to_dev(traj).to(U.device).reshape(-1, U.size(0))
    n, k = x.size(0), U.size(1)
    assert U.size(0) % 2 == 0
    C = torch.empty((k, n), device=x.device)
    check(load().et_to_et_space(ptr(x), n, U.size(0) // 2, ptr(U), k, ptr(C), stream_of(x.device)), "et_to_et_space")
    return back_to(C, traj)


def _to_euclidean_space_raw(C, evec):
    U = to_dev(evec)
    Cd = C.detach()
    if not Cd.is_cuda:
        Cd = Cd.to(U.device)
    if Cd.dtype != torch.float32:
        Cd = Cd.float()
    assert Cd.dim() == 2 and Cd.size(0) == U.size(1)
    n, k = Cd.size(1), Cd.size(0)
    out = torch.empty((n, U.size(0) // 2, 2), device=Cd.device)
    check(load().et_to_euclidean_space(ptr(Cd), Cd.stride(0), Cd.stride(1), n, U.size(0) // 2, ptr(U), k, ptr(out),
                                       stream_of(Cd.device)), "et_to_euclidean_space")
    return back_to(out, C)


class _ToETSpace(torch.autograd.Function):
    """C = U^T M with autograd with respect to the trajectory (``evec`` is detached in the reference, descriptor.py:71):
    the vector-Jacobian product is the other direction of the same pair of kernels, grad_traj = (U grad_C)^T."""

    @staticmethod
    def forward(ctx, traj, evec):
        ctx.save_for_backward(evec)
        ctx.shape = traj.shape
        return _to_et_space_raw(traj, evec)

    @staticmethod
    def backward(ctx, g):
        (evec,) = ctx.saved_tensors
        return _to_euclidean_space_raw(g, evec).reshape(ctx.shape), None


class _ToEuclideanSpace(torch.autograd.Function):
    """traj = (U C)^T with autograd with respect to C (descriptor.py:87): grad_C = U^T grad_traj."""

    @staticmethod
    def forward(ctx, C, evec):
        ctx.save_for_backward(evec)
        return _to_euclidean_space_raw(C, evec)

    @staticmethod
    def backward(ctx, g):
        (evec,) = ctx.saved_tensors
        return _to_et_space_raw(g.contiguous(), evec), None


def to_et_space(traj, evec):
    """C (k,N) = evec^T M,  M = traj.reshape(-1, 2T)^T; differentiable with respect to ``traj``."""
    if _wants_grad(traj):
        return _ToETSpace.apply(traj, evec.detach())
    return _to_et_space_raw(traj, evec)


def to_euclidean_space(C, evec, dim=2):
    """traj (N,T,dim) = (evec C)^T; C (k,N) may be a strided view (e.g. C3[:, :, s]); differentiable with respect to C."""
    assert dim == 2, "only 2-D trajectories are supported"
    if _wants_grad(C):
        return _ToEuclideanSpace.apply(C, evec.detach())
    return _to_euclidean_space_raw(C, evec)


def project(obs, pred, U_obs, U_pred, ori=True, rot=True, sca=True):
    """Fused normalise + projection.  Returns (C_obs, C_pred|None, (ori, rot, sca))."""
    x = to_dev(obs)
    n, t_obs = _ntc(x)
    p = to_dev(pred).to(x.device) if pred is not None else None
    t_pred = p.size(1) if p is not None else 0
    Uo = to_dev(U_obs).to(x.device)
    Up = to_dev(U_pred).to(x.device) if p is not None else None
    k = Uo.size(1)
    C_obs = torch.empty((k, n), device=x.device)
    C_pred = torch.empty((k, n), device=x.device) if p is not None else None
    o = torch.empty((n, 1, 2), device=x.device) if ori else None
    r = torch.empty((n, 2, 2), device=x.device) if rot else None
    s = torch.empty((n, 1, 1), device=x.device) if sca else None
    check(load().et_project(ptr(x), ptr(p), n, t_obs, t_pred, ptr(Uo), ptr(Up), k, norm_flags(ori, rot, sca), ptr(C_obs),
                            ptr(C_pred), ptr(o), ptr(r), ptr(s), stream_of(x.device)), "et_project")
    return (back_to(C_obs, obs), back_to(C_pred, obs), (back_to(o, obs), back_to(r, obs), back_to(s, obs)))


def _reconstruct_raw(C, anchor, U, state, t_pred):
    ori, rot, sca = state
    k, n, s = C.shape
    out = torch.empty((s, n, t_pred, 2), device=C.device)
    flags = norm_flags(ori is not None, rot is not None, sca is not None)
    check(load().et_reconstruct(ptr(C), ptr(anchor), n, s, k, t_pred, ptr(U), flags, ptr(ori), ptr(rot), ptr(sca),
                                ptr(out), stream_of(C.device)), "et_reconstruct")
    return out


class _Reconstruct(torch.autograd.Function):
    """out (S,N,T,2) = denormalise(U (C + anchor)); differentiable wrt C only (descriptor.py:87, anchor.py:87)."""

    @staticmethod
    def forward(ctx, C, anchor, U, ori, rot, sca):
        ctx.save_for_backward(U, rot, sca)
        ctx.flags = norm_flags(ori is not None, rot is not None, sca is not None)
        return _reconstruct_raw(C, anchor, U, (ori, rot, sca), U.size(0) // 2)

    @staticmethod
    def backward(ctx, grad_out):
        U, rot, sca = ctx.saved_tensors
        g = grad_out.contiguous()
        s, n, t, _ = g.shape
        k = U.size(1)
        grad_C = torch.empty((k, n, s), device=g.device)
        check(load().et_reconstruct_bwd(ptr(g), n, s, k, t, ptr(U), ctx.flags, ptr(rot), ptr(sca), ptr(grad_C),
                                        stream_of(g.device)), "et_reconstruct_bwd")
        return grad_C, None, None, None, None, None


def reconstruct(C_pred, U_pred, state, anchor=None):
    """ETDescriptor.reconstruction (optionally fused with ETAnchor.forward).

    C_pred (k,N,S) -> (S,N,T,2); ``state`` = (ori, rot, sca) of the same batch (None = disabled).
    Gradients flow to ``C_pred`` only.
    """
    assert C_pred.dim() == 3
    like = C_pred
    needs_grad = C_pred.requires_grad and torch.is_grad_enabled()
    if not C_pred.is_cuda:
        dev = compute_device()
        Cd = C_pred.to(dev)
    else:
        dev, Cd = C_pred.device, C_pred
    Cd = Cd.float().contiguous()
    U = to_dev(U_pred).to(dev)
    st = tuple(to_dev(v).to(dev) if v is not None else None for v in state)
    n = Cd.size(1)
    for name, v, shape in (("ori", st[0], (n, 1, 2)), ("rot", st[1], (n, 2, 2)), ("sca", st[2], (n, 1, 1))):
        assert v is None or tuple(v.shape) == shape, (
            f"normaliser state {name} has shape {tuple(v.shape)} but the coefficients hold {n} pedestrians")
    a = to_dev(anchor).to(dev) if anchor is not None else None
    if needs_grad:
        out = _Reconstruct.apply(Cd, a, U, *st)
    else:
        out = _reconstruct_raw(Cd.detach(), a, U, st, U.size(0) // 2)
    return back_to(out, like)


def project_reconstruct(obs, pred, U_obs, U_pred, ori=True, rot=True, sca=True, want_coeffs=True, variant=0,
                        out=None):
    """Headline op: rank-k round trip of obs and pred in one pass.

    Returns (rec_obs, rec_pred, C_obs|None, C_pred|None).  ``out`` may supply preallocated device
    tensors (rec_obs, rec_pred, C_obs, C_pred) to keep allocation out of a timed region.
    """
    if not obs.is_cuda and out is None and obs.size(0) >= 2 * HOST_CHUNK:
        return _project_reconstruct_host(obs, pred, U_obs, U_pred, ori, rot, sca, want_coeffs, variant)
    x = to_dev(obs)
    n, t_obs = _ntc(x)
    p = to_dev(pred).to(x.device)
    t_pred = p.size(1)
    Uo, Up = to_dev(U_obs).to(x.device), to_dev(U_pred).to(x.device)
    k = Uo.size(1)
    if out is not None:
        rec_obs, rec_pred, C_obs, C_pred = out
    else:
        rec_obs, rec_pred = torch.empty_like(x), torch.empty_like(p)
        C_obs = torch.empty((k, n), device=x.device) if want_coeffs else None
        C_pred = torch.empty((k, n), device=x.device) if want_coeffs else None
    check(load().et_project_reconstruct(ptr(x), ptr(p), n, t_obs, t_pred, ptr(Uo), ptr(Up), k, norm_flags(ori, rot, sca),
                                        ptr(rec_obs), ptr(rec_pred), ptr(C_obs), ptr(C_pred), variant,
                                        stream_of(x.device)), "et_project_reconstruct")
    return back_to(rec_obs, obs), back_to(rec_pred, obs), back_to(C_obs, obs), back_to(C_pred, obs)


# ----------------------------------------------------------------------------------------
# EigenTrajectory.forward glue without boolean-mask gathers (EigenTrajectory/model.py:73-105)
# ----------------------------------------------------------------------------------------
def forward_project(obs, pred, U_obs_m, U_obs_s, U_pred_m, U_pred_s, static_dist):
    """Moving/static split + both projections in one launch.

    Returns (C_obs (k,N), C_pred (k,N)|None, (ori, rot, sca) of every row, moving (N,) bool), all on the
    compute device.  Static rows carry sca = 1."""
    x = to_dev(obs)
    n, t_obs = _ntc(x)
    dev = x.device
    p = to_dev(pred).to(dev) if pred is not None else None
    t_pred = p.size(1) if p is not None else 0
    Uom, Uos = to_dev(U_obs_m).to(dev), to_dev(U_obs_s).to(dev)
    Upm = to_dev(U_pred_m).to(dev) if p is not None else None
    Ups = to_dev(U_pred_s).to(dev) if p is not None else None
    k = Uom.size(1)
    C_obs = torch.empty((k, n), device=dev)
    C_pred = torch.empty((k, n), device=dev) if p is not None else None
    ori, rot, sca = torch.empty((n, 1, 2), device=dev), torch.empty((n, 2, 2), device=dev), torch.empty((n, 1, 1), device=dev)
    moving = torch.empty((n,), dtype=torch.uint8, device=dev)
    check(load().et_forward_project(ptr(x), ptr(p), n, t_obs, t_pred, ptr(Uom), ptr(Uos), ptr(Upm), ptr(Ups), k,
                                    float(static_dist), ptr(C_obs), ptr(C_pred), ptr(ori), ptr(rot), ptr(sca), ptr(moving),
                                    stream_of(dev)), "et_forward_project")
    return C_obs, C_pred, (ori, rot, sca), moving.view(torch.bool)


def _forward_reconstruct_raw(C, anchor_m, anchor_s, U_m, U_s, moving, ori, rot, sca):
    k, n, s = C.shape
    t = U_m.size(0) // 2
    out = torch.empty((s, n, t, 2), device=C.device)
    check(load().et_forward_reconstruct(ptr(C), ptr(anchor_m), ptr(anchor_s), n, s, k, t, ptr(U_m), ptr(U_s), ptr(moving),
                                        ptr(ori), ptr(rot), ptr(sca), ptr(out), stream_of(C.device)), "et_forward_reconstruct")
    return out


class _ForwardReconstruct(torch.autograd.Function):
    """out (S,N,T,2) = denormalise(U_g (C + anchor_g)), g = moving[n]; differentiable wrt C only."""

    @staticmethod
    def forward(ctx, C, anchor_m, anchor_s, U_m, U_s, moving, ori, rot, sca):
        ctx.save_for_backward(U_m, U_s, moving, rot, sca)
        return _forward_reconstruct_raw(C, anchor_m, anchor_s, U_m, U_s, moving, ori, rot, sca)

    @staticmethod
    def backward(ctx, grad_out):
        U_m, U_s, moving, rot, sca = ctx.saved_tensors
        g = grad_out.contiguous()
        s, n, t, _ = g.shape
        k = U_m.size(1)
        grad_C = torch.empty((k, n, s), device=g.device)
        check(load().et_forward_reconstruct_bwd(ptr(g), n, s, k, t, ptr(U_m), ptr(U_s), ptr(moving), ptr(rot), ptr(sca),
                                                ptr(grad_C), stream_of(g.device)), "et_forward_reconstruct_bwd")
        return (grad_C,) + (None,) * 8


_loss_ws = {}


def _loss_workspace(device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _loss_ws.get(key)
    if ws is None:
        ws = torch.zeros(16, dtype=torch.uint8, device=device)
        _loss_ws[key] = ws
    return ws


class _ForwardReconLosses(torch.autograd.Function):
    """(recon, loss_eigentraj, loss_euclidean_ade, loss_euclidean_fde) of model.py:98-123 in two launches; the backward
    pass is one sparse kernel (only arg-min samples carry gradient) plus the dense reconstruction backward if `recon`
    itself is used downstream."""

    @staticmethod
    def forward(ctx, C, anchor_m, anchor_s, U_m, U_s, moving, ori, rot, sca, C_gt, gt):
        ctx.set_materialize_grads(False)
        recon = _forward_reconstruct_raw(C, anchor_m, anchor_s, U_m, U_s, moving, ori, rot, sca)
        k, n, s = C.shape
        t = recon.size(2)
        per_ped = torch.empty((3, n), device=C.device)
        argmins = torch.empty((3, n), dtype=torch.int32, device=C.device)
        losses = torch.empty((3,), device=C.device)
        check(load().et_forward_losses(ptr(C), ptr(anchor_m), ptr(anchor_s), ptr(moving), ptr(C_gt), ptr(recon), ptr(gt), n, s, k,
                                       t, ptr(per_ped), ptr(argmins), ptr(losses), ptr(_loss_workspace(C.device)),
                                       stream_of(C.device)), "et_forward_losses")
        ctx.save_for_backward(C, anchor_m, anchor_s, U_m, U_s, moving, rot, sca, C_gt, gt, recon, argmins)
        return recon, losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_recon, g_ec, g_ade, g_fde):
        C, anchor_m, anchor_s, U_m, U_s, moving, rot, sca, C_gt, gt, recon, argmins = ctx.saved_tensors
        k, n, s = C.shape
        t = recon.size(2)
        lib = load()
        grad_C = torch.empty((k, n, s), device=C.device)
        accumulate = 0
        if g_recon is not None:
            g = g_recon.contiguous()
            check(lib.et_forward_reconstruct_bwd(ptr(g), n, s, k, t, ptr(U_m), ptr(U_s), ptr(moving), ptr(rot), ptr(sca),
                                                 ptr(grad_C), stream_of(C.device)), "et_forward_reconstruct_bwd")
            accumulate = 1
        if g_ec is not None or g_ade is not None or g_fde is not None:
            zero = torch.zeros((), device=C.device)
            w = torch.stack([zero if v is None else v.float() for v in (g_ec, g_ade, g_fde)])
            check(lib.et_forward_losses_bwd(ptr(C), ptr(anchor_m), ptr(anchor_s), ptr(moving), ptr(C_gt), ptr(recon), ptr(gt), n,
                                            s, k, t, ptr(U_m), ptr(U_s), ptr(rot), ptr(sca), ptr(argmins), ptr(w), accumulate,
                                            ptr(grad_C), stream_of(C.device)), "et_forward_losses_bwd")
        elif not accumulate:
            grad_C.zero_()
        return (grad_C,) + (None,) * 10


def forward_reconstruct_losses(C_pred, anchor_m, anchor_s, U_m, U_s, moving, state, C_pred_gt, pred_gt):
    """Anchor refinement + reconstruction + the three training losses (model.py:98-123).

    Returns (recon (S,N,T,2), loss_eigentraj, loss_euclidean_ade, loss_euclidean_fde); autograd wrt C_pred."""
    dev = state[0].device
    like = C_pred
    Cd = C_pred if C_pred.is_cuda else C_pred.to(dev)
    Cd = Cd.float().contiguous()
    args = (to_dev(anchor_m).to(dev), to_dev(anchor_s).to(dev), to_dev(U_m).to(dev), to_dev(U_s).to(dev),
            moving.view(torch.uint8), *state, to_dev(C_pred_gt).to(dev), to_dev(pred_gt).to(dev))
    recon, l_ec, l_ade, l_fde = _ForwardReconLosses.apply(Cd, *args)
    return back_to(recon, like), back_to(l_ec, like), back_to(l_ade, like), back_to(l_fde, like)


def forward_reconstruct(C_pred, anchor_m, anchor_s, U_m, U_s, moving, state):
    """Anchor refinement + reconstruction of the moving and static rows in one launch (autograd wrt C_pred)."""
    dev = state[0].device
    like = C_pred
    Cd = C_pred if C_pred.is_cuda else C_pred.to(dev)
    Cd = Cd.float().contiguous()
    args = (to_dev(anchor_m).to(dev), to_dev(anchor_s).to(dev), to_dev(U_m).to(dev), to_dev(U_s).to(dev),
            moving.view(torch.uint8), *state)
    if Cd.requires_grad and torch.is_grad_enabled():
        out = _ForwardReconstruct.apply(Cd, *args)
    else:
        out = _forward_reconstruct_raw(Cd.detach(), *args)
    return back_to(out, like)


HOST_CHUNK = 131072          # pedestrians per pipelined chunk of the host-buffer path (multiple of the 128-row tile)
HOST_FIRST_CHUNK = 32768     # a smaller first chunk shortens the pipeline fill
_side_streams = {}


def _streams(device, count=3):
    st = _side_streams.get(device)
    if st is None:
        st = [torch.cuda.Stream(device=device) for _ in range(count)]
        _side_streams[device] = st
    return st


def _host_chunks(n):
    """Cut [0, n) into pipeline chunks: a short first chunk (the D2H direction idles until the first kernel has run)
    and short last chunks (the H2D direction idles while the last results drain), full-size chunks in between."""
    first, tail = min(HOST_FIRST_CHUNK, HOST_CHUNK), [HOST_CHUNK // 2, HOST_CHUNK // 4]
    cuts, a = [], 0
    body_end = (n - sum(tail)) // 128 * 128 if n >= 4 * HOST_CHUNK else n      # chunk starts stay on tile boundaries
    size = first
    while a < body_end:
        b = min(body_end, a + size)
        cuts.append((a, b))
        a, size = b, HOST_CHUNK
    for size in tail[:-1]:
        if n - a > size + tail[-1]:
            cuts.append((a, a + size))
            a += size
    if a < n:
        cuts.append((a, n))          # the last chunk takes the remainder
    return cuts


_numa = {"bound": False, "why": "bind_host_memory_to_device() not called"}


def _read(path):
    with open(path) as f:
        return f.read().strip()


def _cpulist(text):
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_host_memory_to_device(device):
    """Best effort, one process per GPU: run this process on the CPUs of the NUMA node the GPU hangs off and make that
    node the preferred one for its memory (pinned staging buffers included), so that host <-> device copies of several
    ranks do not all cross the same socket link.  A no-op (with the reason recorded) on single-node hosts, when the
    topology is hidden (numa_node = -1) or when the calls are not permitted.  Returns the summary dict."""
    import ctypes
    import glob
    import os
    global _numa
    try:
        nodes = sorted(int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        props = torch.cuda.get_device_properties(device)
        bus = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(_read(f"/sys/bus/pci/devices/{bus}/numa_node"))
        info = {"gpu_pci": bus, "gpu_numa_node": node, "host_numa_nodes": len(nodes)}
        if len(nodes) < 2 or node < 0:
            _numa = dict(info, bound=False, why="single NUMA node visible" if len(nodes) < 2 else "GPU reports no NUMA node")
            return _numa
        cpus = _cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist")) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238          # x86-64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        _numa = dict(info, bound=rc == 0, cpus=len(cpus), why="ok" if rc == 0 else f"set_mempolicy errno {ctypes.get_errno()}")
    except Exception as exc:      # topology files missing, sandboxed syscalls, ...
        _numa = {"bound": False, "why": f"{type(exc).__name__}: {exc}"}
    return _numa


def host_numa_summary():
    return dict(_numa)


def _project_reconstruct_host(obs, pred, U_obs, U_pred, ori, rot, sca, want_coeffs, variant):
    """Host-buffer path of the headline op: the batch is cut into chunks that flow H2D -> kernel -> D2H on three
    streams, so PCIe traffic in both directions overlaps the kernels.  Inputs should be pinned for the copies to be
    asynchronous; results are returned in freshly allocated pinned host tensors."""
    dev = compute_device()
    obs_h = obs.detach().float().contiguous()
    pred_h = pred.detach().float().contiguous()
    n, t_obs = _ntc(obs_h)
    t_pred = pred_h.size(1)
    cur = torch.cuda.current_stream(dev)
    Uo, Up = to_dev(U_obs).to(dev), to_dev(U_pred).to(dev)
    k = Uo.size(1)
    rec_obs = torch.empty((n, t_obs, 2), pin_memory=True)
    rec_pred = torch.empty((n, t_pred, 2), pin_memory=True)
    C_obs = torch.empty((k, n), pin_memory=True) if want_coeffs else None
    C_pred = torch.empty((k, n), pin_memory=True) if want_coeffs else None
    flags = norm_flags(ori, rot, sca)
    streams = _streams(dev)
    lib = load()
    for s in streams:
        s.wait_stream(cur)          # U_* were produced on the caller's stream
    for ci, (a, b) in enumerate(_host_chunks(n)):
        m = b - a
        st = streams[ci % len(streams)]
        sp = C_void_p(st.cuda_stream)
        with torch.cuda.stream(st):
            xo = obs_h[a:b].to(dev, non_blocking=True)
            xp = pred_h[a:b].to(dev, non_blocking=True)
            ro, rp = torch.empty_like(xo), torch.empty_like(xp)
            co = torch.empty((k, m), device=dev) if want_coeffs else None
            cp = torch.empty((k, m), device=dev) if want_coeffs else None
            check(lib.et_project_reconstruct(ptr(xo), ptr(xp), m, t_obs, t_pred, ptr(Uo), ptr(Up), k, flags, ptr(ro), ptr(rp),
                                             ptr(co), ptr(cp), variant, sp), "et_project_reconstruct")
            rec_obs[a:b].copy_(ro, non_blocking=True)
            rec_pred[a:b].copy_(rp, non_blocking=True)
            if want_coeffs:      # (k, m) device block -> columns [a, b) of the (k, N) host matrix: one pitched copy each
                for host, devbuf in ((C_obs, co), (C_pred, cp)):
                    check(lib.et_memcpy_2d_async(C_void_p(host.data_ptr() + 4 * a), 4 * n, ptr(devbuf), 4 * m, 4 * m, k, sp),
                          "et_memcpy_2d_async")
                    devbuf.record_stream(st)
    for s in streams:
        s.synchronize()             # results live in host memory: they must be complete on return
    return rec_obs, rec_pred, C_obs, C_pred


# ----------------------------------------------------------------------------------------
# eigen-basis (ETDescriptor.truncated_SVD)
# ----------------------------------------------------------------------------------------
_gram_ws = {}


def _gram_workspace(device):
    """Zero-initialised scratch of et_gram, one per (device, stream): concurrent streams never share barrier counters."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _gram_ws.get(key)
    if ws is None:
        ws = torch.zeros(int(load().et_gram_workspace_bytes()), dtype=torch.uint8, device=device)
        _gram_ws[key] = ws
    return ws


def gram(obs, pred=None, ori=False, rot=False, sca=False, G_obs=None, G_pred=None):
    """float64 Gram matrices of the (optionally normalised) trajectory matrices, one pass.

    Returns (G_obs (2T_obs,2T_obs), G_pred (2T_pred,2T_pred)|None) on the compute device; pass
    existing ``G_*`` to accumulate into them (row-sharded callers sum these across ranks).
    """
    x = to_dev(obs)
    n, t_obs = _ntc(x)
    p = to_dev(pred).to(x.device) if pred is not None else None
    t_pred = p.size(1) if p is not None else 0
    if G_obs is None:
        G_obs = torch.zeros((2 * t_obs, 2 * t_obs), dtype=torch.float64, device=x.device)
    if p is not None and G_pred is None:
        G_pred = torch.zeros((2 * t_pred, 2 * t_pred), dtype=torch.float64, device=x.device)
    check(load().et_gram(ptr(x), ptr(p), n, t_obs, t_pred, norm_flags(ori, rot, sca), ptr(G_obs), ptr(G_pred),
                         ptr(_gram_workspace(x.device)), stream_of(x.device)), "et_gram")
    return G_obs, (G_pred if p is not None else None)


def gram_init(obs, pred, ori=True, rot=True, sca=True):
    """The data pass of ``ETDescriptor.parameter_initialization`` in one launch: both float64 Gram matrices, the
    normaliser state of every row and the normalised futures.

    Returns (G_obs, G_pred, pred_norm (N,T_pred,2), (ori|None, rot|None, sca|None)) on the compute device."""
    x = to_dev(obs)
    n, t_obs = _ntc(x)
    p = to_dev(pred).to(x.device)
    t_pred = p.size(1)
    dev = x.device
    no, npr = 4 * t_obs * t_obs, 4 * t_pred * t_pred
    G = torch.zeros(no + npr, dtype=torch.float64, device=dev)          # one fill launch for both accumulators
    G_obs, G_pred = G[:no].view(2 * t_obs, 2 * t_obs), G[no:].view(2 * t_pred, 2 * t_pred)
    pred_norm = torch.empty_like(p)
    o = torch.empty((n, 1, 2), device=dev) if ori else None
    r = torch.empty((n, 2, 2), device=dev) if rot else None
    s = torch.empty((n, 1, 1), device=dev) if sca else None
    check(load().et_gram_init(ptr(x), ptr(p), n, t_obs, t_pred, norm_flags(ori, rot, sca), ptr(G_obs), ptr(G_pred),
                              ptr(pred_norm), ptr(o), ptr(r), ptr(s), ptr(_gram_workspace(dev)), stream_of(dev)), "et_gram_init")
    return G_obs, G_pred, pred_norm, (o, r, s)


def eig_basis(G, k, want64=False, info=None):
    """Leading-k eigenpairs of a float64 Gram matrix -> (U (m,k) fp32, S (k) fp32[, U64, S64]).

    ``info``: optional device int32[2] tensor that receives (sweeps, rotations)."""
    assert G.is_cuda and G.dtype == torch.float64 and G.dim() == 2 and G.size(0) == G.size(1)
    G = G.contiguous()
    m = G.size(0)
    U = torch.empty((m, k), device=G.device)
    S = torch.empty((k,), device=G.device)
    U64 = torch.empty((m, k), dtype=torch.float64, device=G.device) if want64 else None
    S64 = torch.empty((k,), dtype=torch.float64, device=G.device) if want64 else None
    check(load().et_eig_jacobi(ptr(G), m, k, ptr(U), ptr(S), ptr(U64), ptr(S64), ptr(info), stream_of(G.device)),
          "et_eig_jacobi")
    return (U, S, U64, S64) if want64 else (U, S)


def eig_basis_pair(G_a, G_b, k):
    """Both bases of a descriptor in one launch -> ((U_a, S_a), (U_b, S_b)); same results as two ``eig_basis`` calls."""
    for G in (G_a, G_b):
        assert G.is_cuda and G.dtype == torch.float64 and G.dim() == 2 and G.size(0) == G.size(1)
    G_a, G_b = G_a.contiguous(), G_b.contiguous()
    ma, mb = G_a.size(0), G_b.size(0)
    U_a, S_a = torch.empty((ma, k), device=G_a.device), torch.empty((k,), device=G_a.device)
    U_b, S_b = torch.empty((mb, k), device=G_a.device), torch.empty((k,), device=G_a.device)
    check(load().et_eig_jacobi_pair(ptr(G_a), ma, ptr(G_b), mb, k, ptr(U_a), ptr(S_a), ptr(U_b), ptr(S_b),
                                    stream_of(G_a.device)), "et_eig_jacobi_pair")
    return (U_a, S_a), (U_b, S_b)


SVD_SMALL_SMEM_BYTES = 200 * 1024


def svd_small_fits(max_rows, t):
    m = 2 * t
    mp = (m + 1) & ~1
    ld = int(max_rows) | 1
    return (mp * ld + mp * mp + mp) * 4 + mp * 4 <= SVD_SMALL_SMEM_BYTES


def svd_small(traj_norm, k, offsets=None):
    """Batched shared-memory one-sided Jacobi SVD of normalised trajectories.

    traj_norm (N,T,2); ``offsets`` (batch+1,) int64 row boundaries (default: one problem over all
    rows).  Returns U (batch, 2T, k), S (batch, k).
    """
    x = to_dev(traj_norm)
    n, t = _ntc(x)
    if offsets is None:
        offsets = torch.tensor([0, n], dtype=torch.int64)
    off_host = offsets.detach().cpu().to(torch.int64)
    batch = off_host.numel() - 1
    sizes = off_host[1:] - off_host[:-1]
    assert batch >= 1 and int(off_host[0]) >= 0 and int(off_host[-1]) <= n and bool((sizes >= 0).all())
    max_rows = int(sizes.max())
    if not svd_small_fits(max_rows, t):
        raise ValueError(f"svd_small: {max_rows} rows x {2 * t} columns do not fit shared memory")
    off = off_host.to(x.device)
    U = torch.empty((batch, 2 * t, k), device=x.device)
    S = torch.empty((batch, k), device=x.device)
    check(load().et_svd_small(ptr(x), ptr(off), batch, max_rows, t, k, ptr(U), ptr(S), stream_of(x.device)),
          "et_svd_small")
    return U, S


# ----------------------------------------------------------------------------------------
# k-means (EigenTrajectory/kmeans.py)
# ----------------------------------------------------------------------------------------
class KMeansWorkspace:
    """Device buffers one BatchKMeans.fit needs; allocated once per (l, d, K, device).

    ``flat`` = [sums (l,d,K) | counts (l,K) | simsum (l)] is ONE contiguous fp64 buffer so that a
    row-sharded caller all-reduces it in place with a single collective."""

    def __init__(self, l, d, k, device, max_iter=0):
        nbytes = int(load().et_kmeans_workspace_bytes(l, d, k))
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        ns, nc = l * d * k, l * k
        self.flat = torch.zeros((ns + nc + l,), dtype=torch.float64, device=device)
        self.sums = self.flat[:ns].view(l, d, k)
        self.counts = self.flat[ns:ns + nc].view(l, k)
        self.simsum = self.flat[ns + nc:]
        self.simsum_last = torch.zeros((l,), dtype=torch.float64, device=device)
        self.err = torch.zeros((1,), dtype=torch.float64, device=device)
        self.status = torch.zeros((2,), dtype=torch.int32, device=device)

    def reset(self):
        self.flat.zero_()
        self.simsum_last.zero_()
        self.status.zero_()


def kmeans_assign(data, centroids, want_labels=True, want_maxsims=True, acc=None, simsum=None, status=None, row_offset=0,
                  n_global=0):
    """get_labels (+ optional accumulation for compute_centroids).  data (l,d,N), centroids (l,d,K) on the GPU.

    ``n_global`` > 0: ``data`` is the row shard [row_offset, row_offset + N) of a data set with ``n_global`` columns
    (same bits as the unsharded call would give for these columns)."""
    l, d, n = data.shape
    k = centroids.size(-1)
    labels = torch.empty((l, n), dtype=torch.int64, device=data.device) if want_labels else None
    maxsims = torch.empty((l, n), dtype=torch.float32, device=data.device) if want_maxsims else None
    sums = counts = ws = None
    if acc is not None:
        sums, counts, ws = acc.sums, acc.counts, acc.ws
    if n_global:
        check(load().et_kmeans_assign_shard(ptr(data), ptr(centroids), l, d, n, k, ptr(labels), ptr(maxsims), ptr(sums),
                                            ptr(counts), ptr(simsum), ptr(ws), ptr(status), int(row_offset), int(n_global),
                                            stream_of(data.device)), "et_kmeans_assign_shard")
    else:
        check(load().et_kmeans_assign(ptr(data), ptr(centroids), l, d, n, k, ptr(labels), ptr(maxsims), ptr(sums), ptr(counts),
                                      ptr(simsum), ptr(ws), ptr(status), stream_of(data.device)), "et_kmeans_assign")
    return maxsims, labels


KMEANS_FUSED_MAX_BATCH = 32     # batch entries l up to which BatchKMeans.fit uses the persistent whole-fit kernel (every entry needs
                                # co-resident blocks; larger batches run one launch pair per iteration)


def kmeans_lloyd(data, centroids, acc, max_iter, tol, want_labels=True):
    """The whole Lloyd loop of ``BatchKMeans.fit`` in one persistent launch (no host sync inside).

    Returns (labels (l,N) int64 of the last assignment | None, centroids (l,d,K) after the last update);
    ``acc.status`` = {converged, iterations}, ``acc.err``, ``acc.simsum_last`` are filled on the device."""
    l, d, n = data.shape
    k = centroids.size(-1)
    out = torch.empty((l, d, k), device=data.device)
    labels = torch.empty((l, n), dtype=torch.int64, device=data.device) if want_labels else None
    check(load().et_kmeans_lloyd(ptr(data), ptr(centroids), l, d, n, k, int(max_iter), float(tol), ptr(out), ptr(labels),
                                 ptr(acc.err), ptr(acc.status), ptr(acc.simsum_last), ptr(acc.ws), stream_of(data.device)),
          "et_kmeans_lloyd")
    return labels, out


def kmeans_lloyd_sharded(data, centroids, acc, max_iter, tol, rank, world, peers, stamp_base, want_labels=True, row_offset=0,
                         n_global=0):
    """Row-sharded whole-fit kernel with the per-iteration exchange over peer memory (see et_kmeans_lloyd_sharded).

    ``peers``: int64 device tensor (world,) of exchange-buffer addresses as mapped in this process."""
    l, d, n = data.shape
    k = centroids.size(-1)
    out = torch.empty((l, d, k), device=centroids.device)
    labels = torch.empty((l, n), dtype=torch.int64, device=centroids.device) if want_labels else None
    check(load().et_kmeans_lloyd_sharded(ptr(data) if n > 0 else None, ptr(centroids), l, d, n, k, int(max_iter), float(tol),
                                         ptr(out), ptr(labels) if n > 0 else None, ptr(acc.err), ptr(acc.status),
                                         ptr(acc.simsum_last), ptr(acc.ws), int(rank), int(world), ptr(peers),
                                         int(stamp_base) & 0xFFFFFFFF, int(row_offset), int(n_global),
                                         stream_of(centroids.device)), "et_kmeans_lloyd_sharded")
    return labels, out


def kmeans_accumulate(data, labels, acc):
    """Masked sums / counts of compute_centroids for given labels (added into ``acc``)."""
    l, d, n = data.shape
    k = acc.sums.size(-1)
    check(load().et_kmeans_accumulate(ptr(data), ptr(labels), l, d, n, k, ptr(acc.sums), ptr(acc.counts), ptr(acc.ws),
                                      stream_of(data.device)), "et_kmeans_accumulate")


def kmeans_finalize(acc, old_centroids, new_centroids, tol=0.0, use_status=False):
    l, d, k = acc.sums.shape
    check(load().et_kmeans_finalize(ptr(acc.sums), ptr(acc.counts), l, d, k, ptr(old_centroids), ptr(new_centroids),
                                    ptr(acc.err), float(tol), ptr(acc.status) if use_status else None,
                                    ptr(acc.simsum), ptr(acc.simsum_last), stream_of(new_centroids.device)),
          "et_kmeans_finalize")


SEED_SCRATCH_EXTRA = 16      # uint64 words behind the l*K candidate keys (barrier counters of the persistent kernel)


def kmeans_farthest_init(data, k, first_index, scratch=None):
    """BatchKMeans.kmeanspp (farthest-point seeding from ``first_index``): all K - 1 steps in one persistent launch."""
    l, d, n = data.shape
    cent = torch.empty((l, d, k), device=data.device)
    if scratch is None:
        scratch = torch.empty((l * k + SEED_SCRATCH_EXTRA,), dtype=torch.int64, device=data.device)
    assert scratch.numel() >= l * k + SEED_SCRATCH_EXTRA
    check(load().et_kmeans_farthest_init(ptr(data), l, d, n, k, int(first_index), ptr(cent), ptr(scratch),
                                         stream_of(data.device)), "et_kmeans_farthest_init")
    return cent


def kmeans_d2_init(data, k, uniform, trials=None):
    """k-means++ seeding by D^2 sampling with greedy local trials (sklearn's ``init="k-means++"``), one persistent launch.

    ``uniform``: (l, K, trials) float64 numbers in [0, 1) from the caller's generator (host or device)."""
    l, d, n = data.shape
    u = torch.as_tensor(uniform, dtype=torch.float64).to(data.device).contiguous()
    if trials is None:
        trials = u.size(-1)
    assert tuple(u.shape) == (l, k, trials), f"uniform has shape {tuple(u.shape)}, expected {(l, k, trials)}"
    cent = torch.empty((l, d, k), device=data.device)
    ws = torch.empty(int(load().et_kmeans_d2_workspace_bytes(l, trials)), dtype=torch.uint8, device=data.device)
    check(load().et_kmeans_d2_init(ptr(data), l, d, n, k, int(trials), ptr(u), ptr(cent), ptr(ws), stream_of(data.device)),
          "et_kmeans_d2_init")
    return cent


def kmeans_farthest_init_sharded(data, k, first_global_index, row_offset, n_global, rank, world, peers, stamp_base):
    """Farthest-point seeding over row shards with the per-step candidate exchange inside the persistent kernel (peer
    memory, see et_kmeans_farthest_init_sharded).  Returns the (l,d,K) centroids, identical on every rank."""
    l, d, n = data.shape
    dev = peers.device
    cent = torch.empty((l, d, k), device=dev)
    scratch = torch.empty((l * k + SEED_SCRATCH_EXTRA,), dtype=torch.int64, device=dev)
    check(load().et_kmeans_farthest_init_sharded(ptr(data) if n > 0 else None, l, d, n, k, int(first_global_index),
                                                 int(row_offset), int(n_global), ptr(cent), ptr(scratch), int(rank), int(world),
                                                 ptr(peers), int(stamp_base) & 0xFFFFFFFF, stream_of(dev)),
          "et_kmeans_farthest_init_sharded")
    return cent


def kmeans_seed_step(data, centroids, ncols):
    """Local farthest-point candidate of a row shard: returns (best similarity (l,) fp32, local index (l,) int64)."""
    l, d, n = data.shape
    k = centroids.size(-1)
    key = torch.empty((l,), dtype=torch.int64, device=data.device)
    check(load().et_kmeans_seed_step(ptr(data), ptr(centroids), l, d, n, k, int(ncols), ptr(key), stream_of(data.device)),
          "et_kmeans_seed_step")
    idx = key & 0xFFFFFFFF
    bits = (key >> 32) & 0xFFFFFFFF
    # undo the order-preserving map: keys with the top bit set were non-negative floats
    raw = torch.where(bits >= 0x80000000, bits - 0x80000000, 0xFFFFFFFF - bits)
    return raw.to(torch.uint32).view(torch.float32), idx


def kmeans_seed_candidate(data, centroids, ncols, row_offset, out=None, n_global=0):
    """Global, signed-orderable candidate key (l,) int64 of a row shard (see et_kmeans_seed_candidate)."""
    l, d, n = data.shape
    k = centroids.size(-1)
    key = out if out is not None else torch.empty((l,), dtype=torch.int64, device=data.device)
    check(load().et_kmeans_seed_candidate(ptr(data), ptr(centroids), l, d, n, k, int(ncols), int(row_offset), int(n_global), ptr(key),
                                          stream_of(data.device)), "et_kmeans_seed_candidate")
    return key


def kmeans_seed_fetch(data, row_offset, gkey, out=None):
    """(l,d) float64 coordinates of the elected column on the owning rank, zeros elsewhere."""
    l, d, n = data.shape
    coords = out if out is not None else torch.empty((l, d), dtype=torch.float64, device=data.device)
    check(load().et_kmeans_seed_fetch(ptr(data), l, d, n, int(row_offset), ptr(gkey), ptr(coords), stream_of(data.device)),
          "et_kmeans_seed_fetch")
    return coords


# ----------------------------------------------------------------------------------------
# metrics (utils/metrics.py)
# ----------------------------------------------------------------------------------------
def ade_fde(pred, gt, want_argmin=False, want_tcc=False):
    """min-over-samples ADE and FDE per pedestrian in one pass (optionally + arg-min of FDE and the TCC).

    pred (S,N,T,2), gt (N,T,2)|(1,N,T,2) -> (ade, fde[, argmin][, tcc])."""
    p = to_dev(pred)
    g = to_dev(gt).to(p.device)
    if g.dim() == 4:
        assert g.size(0) == 1
        g = g[0]
    s, n, t, c = p.shape
    assert c == 2 and tuple(g.shape) == (n, t, 2), f"pred {tuple(p.shape)} vs gt {tuple(g.shape)}"
    ade = torch.empty((n,), device=p.device)
    fde = torch.empty((n,), device=p.device)
    arg = torch.empty((n,), dtype=torch.int32, device=p.device) if want_argmin else None
    tcc = torch.empty((n,), device=p.device) if want_tcc else None
    check(load().et_ade_fde(ptr(p), ptr(g), s, n, t, ptr(ade), ptr(fde), ptr(arg), ptr(tcc), stream_of(p.device)),
          "et_ade_fde")
    out = [ade, fde]
    if want_argmin:
        out.append(arg)
    if want_tcc:
        out.append(tcc)
    return tuple(out)


def col(pred, thres=0.2):
    """Collision rate per pedestrian of ONE scene: pred (S,N,T,2) -> (N,) percent."""
    p = to_dev(pred)
    s, n, t, c = p.shape
    assert c == 2
    out = torch.empty((n,), device=p.device)
    check(load().et_col(ptr(p), s, n, t, float(thres), ptr(out), stream_of(p.device)), "et_col")
    return out
