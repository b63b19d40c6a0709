import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def t(a):
    """numpy -> CPU torch tensor"""
    return torch.from_numpy(np.ascontiguousarray(a))


def rel_max(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def rel_fro(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def align_signs(U, U_ref):
    """Flip columns of U so that each has a positive inner product with the reference column."""
    U, U_ref = torch.as_tensor(U).double(), torch.as_tensor(U_ref).double()
    s = torch.sign((U * U_ref).sum(dim=0))
    s[s == 0] = 1
    return U * s, s


def spawn_ranks(worker, world, extra_args, out=None, attempts=3):
    """``mp.spawn(worker, (world, port, *extra_args))`` on a free localhost port.  A run that fails BEFORE all ranks have
    found each other (the worker sets ``out["ready", rank]`` right after ``init_process_group``) or with a network-level
    message (port taken between probing and binding, connection reset) is infrastructure, not the code under test, and
    is retried on a fresh port; every other failure propagates unchanged."""
    import socket
    import torch.multiprocessing as mp
    network = ("address already in use", "eaddrinuse", "connection refused", "connection reset", "timed out", "broken pipe",
               "socket", "connect() ")
    for attempt in range(attempts):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        try:
            mp.spawn(worker, args=(world, port, *extra_args), nprocs=world, join=True)
            return
        except Exception as exc:      # noqa: BLE001 - inspected and re-raised below
            text = str(exc).lower()
            met = out is not None and all(out.get(("ready", r)) for r in range(world))
            if attempt + 1 < attempts and (any(k in text for k in network) or (out is not None and not met)):
                print(f"rendezvous failed on port {port} ({type(exc).__name__}: {text[:200]}); retrying on a new port")
                if out is not None:
                    out.clear()
                continue
            raise
