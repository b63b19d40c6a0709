"""Context for the roofline fractions: what plain streaming kernels reach on this GPU at the sizes of our ops
(pure write, pure read, copy), L2 flushed between launches."""
import os, sys, json, torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(root, "scripts"))
from bench_kernels import time_op
dev = torch.device("cuda")
out = []
for mb in (160, 400, 2000):
    n = mb * 1000 * 1000 // 4
    x = torch.empty(n, device=dev); y = torch.empty(n, device=dev); x.normal_()
    a, m = time_op(lambda: y.zero_(), reps=20)
    out.append({"op": "write (fill)", "MB": mb, "us": 1e3 * a, "GBs": mb * 1e6 / (a * 1e-3) / 1e9})
    a, m = time_op(lambda: x.sum(), reps=20)
    out.append({"op": "read (sum)", "MB": mb, "us": 1e3 * a, "GBs": mb * 1e6 / (a * 1e-3) / 1e9})
    a, m = time_op(lambda: y.copy_(x), reps=20)
    out.append({"op": "copy (read+write)", "MB": 2 * mb, "us": 1e3 * a, "GBs": 2 * mb * 1e6 / (a * 1e-3) / 1e9})
    del x, y
for r in out:
    print(json.dumps(r))
json.dump(out, open(os.path.join(root, "gpurun_out", "hbm_ceilings.json"), "w"), indent=1)
