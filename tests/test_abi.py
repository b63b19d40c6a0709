"""CPU: the C-ABI library loads and exports every symbol include/et_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "et_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(et_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from eigentrajectory_b200 import _lib
    return _lib


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("et_project_reconstruct", "et_project", "et_reconstruct", "et_reconstruct_bwd", "et_gram",
                 "et_eig_jacobi", "et_svd_small", "et_kmeans_assign", "et_kmeans_finalize", "et_kmeans_farthest_init",
                 "et_ade_fde", "et_norm_params", "et_normalize", "et_denormalize"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    cdll = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, f"declared in et_b200.h but not exported: {missing}"


def test_python_prototypes_cover_the_header(lib):
    assert sorted(lib.PROTOTYPES) == declared_symbols()
    handle = lib.load()
    assert handle.et_version() == 100
    assert handle.et_launch_count() == 0          # nothing has been launched on a CPU-only host


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import eigentrajectory_b200 as et
    with pytest.raises(et.ETLibraryError):
        et.TrajNorm().calculate_params(torch.zeros(4, 8, 2))
    with pytest.raises(et.ETLibraryError):
        et.compute_batch_ade(torch.zeros(20, 4, 12, 2), torch.zeros(4, 12, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eigentrajectory_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"


def test_synthetic_generators_agree():
    """The product's benchmark generator and the oracle's copy produce identical bits."""
    import torch
    from eigentrajectory_b200.synthetic import synthetic_trajectories as mine
    from oracle.et_oracle import synthetic_trajectories as theirs
    for n, seed in ((1000, 0), (257, 1003)):
        a, b = mine(n, seed), theirs(n, seed)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
