"""ctypes binding of ``libet_b200.so`` (the C ABI declared in ``include/et_b200.h``).

There is no CPU or PyTorch fallback: if the shared library is missing, cannot be loaded, or a
call fails, an exception is raised.  Build it with ``make`` at the repository root (or
``python -c "import __graft_entry__ as g; g.build()"``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libet_b200.so")

ET_NORM_ORI, ET_NORM_ROT, ET_NORM_SCA = 1, 2, 4
ET_MAX_T, ET_MAX_K, ET_MAX_CLUSTERS, ET_MAX_KM_DIM = 32, 32, 64, 16

_p, _i, _l, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/et_b200.h one to one
PROTOTYPES = {
    "et_version": (_i, []),
    "et_last_error": (C.c_char_p, []),
    "et_device_info": (_i, [_p, _p, _p]),
    "et_launch_count": (_l, []),
    "et_tune": (_i, [_i, _i]),
    "et_memcpy_2d_async": (_i, [_p, _sz, _p, _sz, _sz, _sz, _p]),
    "et_norm_params": (_i, [_p, _l, _i, _i, _p, _p, _p, _p]),
    "et_normalize": (_i, [_p, _l, _i, _i, _p, _p, _p, _p, _p]),
    "et_denormalize": (_i, [_p, _l, _i, _i, _p, _p, _p, _p, _p]),
    "et_to_et_space": (_i, [_p, _l, _i, _p, _i, _p, _p]),
    "et_to_euclidean_space": (_i, [_p, _l, _l, _l, _i, _p, _i, _p, _p]),
    "et_project": (_i, [_p, _p, _l, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "et_reconstruct": (_i, [_p, _p, _l, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "et_reconstruct_bwd": (_i, [_p, _l, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "et_project_reconstruct": (_i, [_p, _p, _l, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _i, _p]),
    "et_forward_project": (_i, [_p, _p, _l, _i, _i, _p, _p, _p, _p, _i, C.c_float, _p, _p, _p, _p, _p, _p, _p]),
    "et_forward_reconstruct": (_i, [_p, _p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "et_forward_reconstruct_bwd": (_i, [_p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "et_forward_losses": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p]),
    "et_forward_losses_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "et_gram_workspace_bytes": (_sz, []),
    "et_gram": (_i, [_p, _p, _l, _i, _i, _i, _p, _p, _p, _p]),
    "et_gram_init": (_i, [_p, _p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "et_eig_jacobi": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "et_eig_jacobi_pair": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _p, _p]),
    "et_svd_small": (_i, [_p, _p, _i, _l, _i, _i, _p, _p, _p]),
    "et_kmeans_workspace_bytes": (_sz, [_i, _i, _i]),
    "et_kmeans_assign": (_i, [_p, _p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "et_kmeans_assign_shard": (_i, [_p, _p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _p, _p, _l, _l, _p]),
    "et_kmeans_lloyd": (_i, [_p, _p, _i, _i, _l, _i, _i, _d, _p, _p, _p, _p, _p, _p, _p]),
    "et_kmeans_exchange_bytes": (_sz, [_i, _i, _i, _i]),
    "et_kmeans_lloyd_sharded": (_i, [_p, _p, _i, _i, _l, _i, _i, _d, _p, _p, _p, _p, _p, _p, _i, _i, _p, C.c_uint, _l, _l, _p]),
    "et_kmeans_accumulate": (_i, [_p, _p, _i, _i, _l, _i, _p, _p, _p, _p]),
    "et_kmeans_finalize": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _d, _p, _p, _p, _p]),
    "et_kmeans_farthest_init": (_i, [_p, _i, _i, _l, _i, _l, _p, _p, _p]),
    "et_kmeans_d2_workspace_bytes": (_sz, [_i, _i]),
    "et_kmeans_d2_init": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _p, _p]),
    "et_kmeans_farthest_init_sharded": (_i, [_p, _i, _i, _l, _i, _l, _l, _l, _p, _p, _i, _i, _p, C.c_uint, _p]),
    "et_kmeans_seed_step": (_i, [_p, _p, _i, _i, _l, _i, _i, _p, _p]),
    "et_kmeans_seed_candidate": (_i, [_p, _p, _i, _i, _l, _i, _i, _l, _l, _p, _p]),
    "et_kmeans_seed_fetch": (_i, [_p, _i, _i, _l, _l, _p, _p, _p]),
    "et_comm_unique_id": (_i, [_p]),
    "et_comm_init": (_i, [_i, _i, _p, _p]),
    "et_comm_rank": (_i, [_p, _p, _p]),
    "et_allreduce_f64": (_i, [_p, _sz, _p, _p]),
    "et_allreduce_min_i64": (_i, [_p, _sz, _p, _p]),
    "et_comm_destroy": (_i, [_p]),
    "et_ade_fde": (_i, [_p, _p, _i, _l, _i, _p, _p, _p, _p, _p]),
    "et_col": (_i, [_p, _i, _l, _i, C.c_float, _p, _p]),
    "et_dataset_parse_host": (_i, [C.c_char_p, _sz, C.c_char, _p, _l, _p]),
    "et_dataset_windows_host": (_i, [_p, _l, _i, _i, _i, _d, _i, _p, _p, _p, _l, _l, _p, _p]),
}

_lib = None


class ETLibraryError(RuntimeError):
    """The native library is missing / unloadable, or one of its entry points failed."""


def load():
    """Load ``libet_b200.so`` once and attach the prototypes.  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ETLibraryError(
            f"{LIB_PATH} not found: the CUDA library has not been built (run `make` in the repository "
            "root).  eigentrajectory_b200 has no CPU / PyTorch fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise ETLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise ETLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise ETLibraryError("eigentrajectory_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


_guard = threading.local()


def check(rc, what):
    """Raise on a non-zero return code.  Also ends the device guard opened by :func:`stream_of` for this call."""
    prev = getattr(_guard, "prev", None)
    if prev is not None:
        _guard.prev = None
        if prev >= 0:
            torch.cuda.set_device(prev)
    if rc != 0:
        msg = load().et_last_error().decode("utf-8", "replace")
        raise ETLibraryError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(device):
    """Current stream of ``device`` as a ``cudaStream_t`` -- and makes ``device`` the CURRENT CUDA device until the
    matching :func:`check` (every call site reads ``check(lib.fn(..., stream_of(dev)), name)``: the arguments are
    evaluated before the call, ``check`` runs after it).  The library sizes its grids, sets kernel attributes and
    launches (cooperatively) on the current device, so a tensor on ``cuda:1`` must not be processed while ``cuda:0``
    is current; the caller's current device is restored afterwards."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cur = torch.cuda.current_device()
    if idx != cur:
        torch.cuda.set_device(idx)
        _guard.prev = cur
    else:
        _guard.prev = -1
    return C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def launch_count():
    return int(load().et_launch_count())
