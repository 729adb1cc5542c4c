#!/usr/bin/env python
"""Development probe (GPU box): one pair x N samples (a cfg 4 share), device time by CUDA events."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
one = torch.from_numpy(np.ascontiguousarray(wl.cfg2_pair()).view(np.float32)).cuda()
d_h = torch.zeros(1, dtype=torch.int64, device="cuda")
for n in [int(float(a)) for a in sys.argv[1:]] or [12_500_000_000]:
    chunk, n_chunks = ctx.plan_debug(0, 1, n)
    ts = []
    for r in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctx.count_fused(one, 1, n, 7, d_h); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"1 pair x {n:.3e}: best {min(ts[1:]):9.3f} ms  {n / min(ts[1:]) / 1e6:8.2f} Gtests/s   chunk {chunk} x {n_chunks} items  hits {int(d_h.item())}")
