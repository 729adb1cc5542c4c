#!/usr/bin/env python
"""Small invocation of every entry point, for `compute-sanitizer --tool memcheck python tools/sanitize_run.py`."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
ctx = satmc.Context(0)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.float32) if a.dtype.fields else np.ascontiguousarray(a)).cuda()
for sv in (False, True):
    pairs = wl.dataset_pairs(300, 3, shape_variance=sv)
    h = ctx.count_fused_host(pairs, 777, 5, sample_offset=3)
    ndof = 5 if sv else 3
    for n in (1, 127, 128, 1000, 4099):
        z = wl.normal_bank(n, ndof, seed=n)
        ctx.count_streamed_host(pairs[:37], z)
    z = wl.normal_bank(37 * 640, ndof, seed=1)
    ctx.count_streamed_host(pairs[:37], z, n_samples=640, z_pair_stride=640)
    one = pairs[:1]
    ctx.count_fused_host(one, 3_000_001, 9)
    d_z = torch.zeros(ndof * 1001, device="cuda"); ctx.fused_normals(1, 2, 7, 1001, ndof, d_z, 1001)
    d_out = torch.zeros(1001, dtype=torch.uint8, device="cuda"); ctx.decide_streamed(put(one), d_z, 1001, ndof, 1001, d_out)
r1, r2 = wl.cfg1_rect_pairs(1000, 1)
d_o = torch.zeros(1000, dtype=torch.uint8, device="cuda"); ctx.sat_corners(put(r1.ravel()), put(r2.ravel()), 1000, d_o)
pairs = wl.dataset_pairs(500, 4)
rb, poses, sds, pi, si, pos = wl.reference_tables(pairs)
bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-3, 4e-3, 8e-3], np.float32)
d_cp = torch.zeros(500, device="cuda"); d_ns = torch.zeros(500, dtype=torch.int32, device="cuda")
it, drawn = ctx.adaptive_run(put(rb), put(poses.ravel()), 500, put(sds.ravel()), 500, put(pi), put(si), put(pos.ravel()), 500,
                             put(bins), put(acc), 4, 30000, 1000, 4000, 5000, 3, d_cp, d_ns)
d_pos = torch.zeros(1000, device="cuda"); d_a = torch.zeros(500, device="cuda"); d_b = torch.zeros(500, device="cuda")
ctx.sample_positions(put(poses.ravel()), 500, put(sds.ravel()), 500, 500, 1.45, 4.0, 1, d_pos, d_a, d_b)
ctx.synchronize()
print("sanitize_run ok", int(h.sum()), it, drawn, float(d_cp.mean()))
