"""GPU (two or more devices): the collective entry points of the C ABI (et_comm_init / et_allreduce_f64 /
et_allreduce_min_i64) driving a row-sharded eigen-basis exactly as a non-Python host would: local et_gram ->
et_allreduce_f64 -> et_eig_jacobi_pair, compared with the unsharded basis."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL ranks)")]


def _worker(rank, world, uid_path, out):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import eigentrajectory_b200 as et
    from eigentrajectory_b200 import ops
    from eigentrajectory_b200._lib import check
    from oracle import et_oracle as O
    lib = et.load_library()
    uid = (C.c_char * 128)()
    if rank == 0:
        check(lib.et_comm_unique_id(uid), "et_comm_unique_id")
        with open(uid_path + ".tmp", "wb") as f:
            f.write(bytes(uid))
        os.replace(uid_path + ".tmp", uid_path)
    else:
        import time
        while not os.path.exists(uid_path):
            time.sleep(0.01)
        uid = (C.c_char * 128).from_buffer_copy(open(uid_path, "rb").read())
    comm = C.c_void_p()
    check(lib.et_comm_init(rank, world, uid, C.byref(comm)), "et_comm_init")
    r, n = C.c_int(), C.c_int()
    check(lib.et_comm_rank(comm, C.byref(r), C.byref(n)), "et_comm_rank")
    assert (r.value, n.value) == (rank, world)
    total = 100_001
    obs, pred = O.synthetic_trajectories(total, seed=21)
    a, b = rank * total // world, (rank + 1) * total // world
    G_o, G_p = ops.gram(obs[a:b].to(dev), pred[a:b].to(dev), True, True, True)
    packed = torch.cat([G_o.reshape(-1), G_p.reshape(-1)]).contiguous()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.et_allreduce_f64(C.c_void_p(packed.data_ptr()), packed.numel(), comm, stream), "et_allreduce_f64")
    key = torch.tensor([1000 - rank, rank], dtype=torch.int64, device=dev)
    check(lib.et_allreduce_min_i64(C.c_void_p(key.data_ptr()), 2, comm, stream), "et_allreduce_min_i64")
    (Uo, So), (Up, Sp) = ops.eig_basis_pair(packed[:256].reshape(16, 16).contiguous(), packed[256:].reshape(24, 24).contiguous(), 6)
    F_o, F_p = ops.gram(obs.to(dev), pred.to(dev), True, True, True)
    (Uo1, _), (Up1, Sp1) = ops.eig_basis_pair(F_o, F_p, 6)
    out[rank] = dict(key=key.tolist(), gram_rel=float((packed[256:] - F_p.reshape(-1)).abs().max() / F_p.abs().max()),
                     proj=float((Up.double() @ Up.double().T - Up1.double() @ Up1.double().T).norm()), U=Up.cpu())
    torch.cuda.synchronize()
    check(lib.et_comm_destroy(comm), "et_comm_destroy")


def test_c_abi_collectives_sharded_basis(tmp_path):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, str(tmp_path / "uid"), out), nprocs=world, join=True)
    for r in range(world):
        assert out[r]["key"] == [1000 - (world - 1), 0]
        assert out[r]["gram_rel"] < 1e-12 and out[r]["proj"] < 1e-6
    assert torch.equal(out[0]["U"], out[1]["U"])
