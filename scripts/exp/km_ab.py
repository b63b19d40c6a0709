"""A/B: whole-fit k-means kernel of this build vs the round-1 library, raw ctypes calls.

The round-1 library is not kept in the tree; build it next to this script first:
    mkdir -p /tmp/r01 && git archive 5235d0c | tar -x -C /tmp/r01 && make -C /tmp/r01 eigentrajectory_b200/libet_b200.so
    mkdir -p scripts/exp/old && cp /tmp/r01/eigentrajectory_b200/libet_b200.so scripts/exp/old/libet_b200_r01.so
Measured (B200, us per Lloyd iteration, r01 / single barrier + every block folds): gauss 1e6 21.56 / 21.31, C_pred 1e6
21.57 / 21.46, C_pred 1.25e6 24.28 / 23.97; with the fold shared by groups of four: 19.6 (1e6), 21.9 (1.25e6)."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda")
new = et.load_library()
old = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "old", "libet_b200_r01.so"))
P, I, L, D = C.c_void_p, C.c_int, C.c_int64, C.c_double
old.et_kmeans_lloyd.restype = I
old.et_kmeans_lloyd.argtypes = [P, P, I, I, L, I, I, D, P, P, P, P, P, P, P]
old.et_kmeans_workspace_bytes.restype = C.c_size_t
old.et_kmeans_workspace_bytes.argtypes = [I, I, I]

def datasets():
    gen = torch.Generator().manual_seed(1234)
    g = (torch.randn(1, 6, 1_000_000, generator=gen) * torch.tensor([20., 4., 1., .8, .3, .25])[None, :, None]).contiguous().to(dev)
    yield "gauss 1e6", g
    for n in (1_000_000, 1_250_000):
        obs, pred = (x.to(dev) for x in synthetic_trajectories(n, seed=1000))
        d = et.ETDescriptor(et.DotDict(obs_len=8, pred_len=12, k=6, num_samples=20, traj_dim=2, static_dist=0.3)).to(dev)
        d.parameter_initialization(obs, pred)
        yield f"C_pred {n}", d.projection(obs, pred)[1].unsqueeze(0).contiguous()

for name, data in datasets():
    n = data.size(-1)
    cent = ops.kmeans_farthest_init(data, 20, 12345)
    for tag, lib in (("r01", old), ("now", new)):
        ws = torch.zeros(int(lib.et_kmeans_workspace_bytes(1, 6, 20)), dtype=torch.uint8, device=dev)
        out = torch.empty_like(cent); err = torch.zeros(1, dtype=torch.float64, device=dev)
        status = torch.zeros(2, dtype=torch.int32, device=dev); sims = torch.zeros(1, dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        def run(iters):
            rc = lib.et_kmeans_lloyd(C.c_void_p(data.data_ptr()), C.c_void_p(cent.data_ptr()), 1, 6, n, 20, iters, -1.0,
                                     C.c_void_p(out.data_ptr()), None, C.c_void_p(err.data_ptr()), C.c_void_p(status.data_ptr()),
                                     C.c_void_p(sims.data_ptr()), C.c_void_p(ws.data_ptr()), st)
            assert rc == 0
        for _ in range(2): run(100)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(100); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 10)
        print(f"{name:16s} {tag}: us per Lloyd iteration {min(ts):.2f}  nan centroids {int(out.isnan().sum())}")
