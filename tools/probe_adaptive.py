#!/usr/bin/env python
"""Development probe (GPU box): wall time of one adaptive z-test batch (bench.py's adaptive_batch), several repeats."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
if os.environ.get("SATMC_LIB"):
    satmc.LIB_PATH = os.environ["SATMC_LIB"]
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
pairs = wl.dataset_pairs(100_000, 3)
rb, poses, sds, pi, si, pos = wl.reference_tables(pairs)
bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-4, 1e-3, 1e-2], np.float32)
dd = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rb, poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
d_hits = torch.zeros(pairs.size, device="cuda")
ts = []
for r in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    it, drawn = ctx.adaptive_run(dd[0], dd[1], pairs.size, dd[2], pairs.size, dd[3], dd[4], dd[5], pairs.size, dd[6], dd[7], 4,
                                 1_020_000, 1000, 20000, 100000, 7, d_hits)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(f"adaptive batch: min {min(ts):.2f} ms, median {sorted(ts)[len(ts) // 2]:.2f} ms, iterations {it}, samples {drawn}, sum {float(d_hits.sum()):.3f}")
