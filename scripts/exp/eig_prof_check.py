"""Do the instrumented variants of the eigen-solve (ET_TUNE_EIG_THREADS 3001 / 3002) run as long as the plain ones (2001 / 2002)?
Event-timed next to each other, with the instrumented body's own entry-to-exit clock64 ticks and global-timer nanoseconds."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import eigentrajectory_b200 as et
from eigentrajectory_b200 import ops
from eigentrajectory_b200.synthetic import synthetic_trajectories
dev = torch.device("cuda"); lib = et.load_library()
obs, pred = (x.to(dev) for x in synthetic_trajectories(1_000_000, seed=0))
G_o, G_p, _, _ = ops.gram_init(obs, pred)
U = torch.empty((24, 6), device=dev); S = torch.empty(6, device=dev)
info = torch.zeros(14, dtype=torch.int32, device=dev)
from eigentrajectory_b200._lib import ptr, stream_of
def raw():      # the bare C call: no allocations in the timed region
    lib.et_eig_jacobi(ptr(G_p), 24, 6, ptr(U), ptr(S), None, None, ptr(info), stream_of(dev))
def timed(fn, reps=40):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return round(ts[len(ts) // 2], 1), round(ts[0], 1)
def back_to_back(fn, n=50):     # n launches between one event pair: launch overhead hidden behind the previous kernel
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / n, 1)
for knob in (2001, 3001, 2002, 3002):
    lib.et_tune(3, knob)
    med, mn = timed(raw)
    b2b = back_to_back(raw)
    v = info.tolist()
    print(json.dumps({"knob": knob, "single_call_us_median_min": [med, mn], "back_to_back_us_per_call": b2b,
                      "entry_to_exit_ticks": v[12] if knob > 3000 else None, "entry_to_exit_ns": v[13] if knob > 3000 else None,
                      "phase_sum_ticks": sum(v[2:7]) + sum(v[9:12]) if knob > 3000 else None}), flush=True)
lib.et_tune(3, 0)
print(json.dumps({"pair_back_to_back_us": back_to_back(lambda: ops.eig_basis_pair(G_o, G_p, 6))}))
