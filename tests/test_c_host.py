"""The C ABI from a plain-C host (examples/c_host.c): the header is valid C99, a program that includes nothing but it and
the CUDA runtime links against libet_b200.so, and -- on a GPU -- computes the eigen-basis and the rank-6 round trip of
synthetic pedestrians (ETDescriptor.parameter_initialization + projection / reconstruction, descriptor.py:116-176)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA_HOME, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not installed")
    pkg = os.path.join(ROOT, "eigentrajectory_b200")
    assert os.path.exists(os.path.join(pkg, "libet_b200.so")), "build the library first (make)"
    exe = str(tmp_path / "c_host")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror=implicit-function-declaration", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(CUDA_HOME, "include"), os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L" + pkg, "-let_b200",
           "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + pkg]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_host_compiles_links_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "libet_b200 version 100" in res.stdout
    if not torch.cuda.is_available():
        assert res.returncode == 2 and "no CPU path" in res.stderr, (res.returncode, res.stderr)


@pytest.mark.gpu
def test_c_host_round_trip_on_the_gpu(tmp_path):
    exe = build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(res.stdout)
    # measured [B200]: relative error 2.0e-3, three kernels (Gram pass, both eigen-solves, fused round trip)
    assert res.returncode == 0, (res.stdout, res.stderr)
    assert "(3 kernels launched)" in res.stdout
