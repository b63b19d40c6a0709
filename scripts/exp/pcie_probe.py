"""Raw host <-> device copy ceiling with N ranks copying at once (torchrun): per rank 160 MB H2D and 208 MB D2H from / to
pinned memory, (a) one direction at a time, (b) both directions concurrently on two streams -- the ceiling the
host-buffer path of bench.py's `e2e` can reach on this host at N GPUs.  One JSON line on rank 0."""
import json, os, time, torch, torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
h_in, h_out = torch.empty(160_000_000, dtype=torch.uint8).pin_memory(), torch.empty(208_000_000, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(160_000_000, dtype=torch.uint8, device=dev), torch.empty(208_000_000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=10):
    ts = []
    for _ in range(reps + 2):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([sum(ts[2:]) / reps], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
res = {"n_gpus": world}
for mode, nbytes in (("h2d", 160e6), ("d2h", 208e6), ("both", 368e6)):
    t = run(mode)
    res[mode + "_ms"] = 1e3 * t; res[mode + "_gbs_per_gpu"] = nbytes / t / 1e9; res[mode + "_gbs_total"] = world * nbytes / t / 1e9
res["e2e_ceiling_traj_per_s"] = world * 1e6 / (res["both_ms"] * 1e-3)
if rank == 0: print(json.dumps(res), flush=True)
if world > 1: dist.destroy_process_group()
