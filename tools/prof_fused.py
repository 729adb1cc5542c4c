#!/usr/bin/env python
"""Short driver for ncu: a few launches of one counting kernel.
usage: prof_fused.py [fused|fused5|streamed3|streamed5|ztest|sweep|poly|cfg4|cfg2|adaptive]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.float32)).cuda()
if mode == "sweep":
    base = wl.dataset_pairs(10_000, 5)
    grid = np.array([0.01, 0.05, 0.15, 0.3]); vx, vy, vt = np.meshgrid(grid, grid, grid, indexing="ij")
    sig = np.sqrt(np.stack([vx.ravel(), vy.ravel(), vt.ravel()], 1)).astype(np.float32)
    d_pairs = put(base); d_s = sig; d_hits = torch.zeros(base.size * 64, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.count_fused_sweep(d_pairs, base.size, d_s, 64, 20_000, 7, d_hits)
elif mode == "poly":
    pr = wl.dataset_pairs(100_000, 3)
    rv = lambda w, h: np.stack([-w / 2, -h / 2, w / 2, -h / 2, w / 2, h / 2, -w / 2, h / 2], 1).reshape(-1, 4, 2).astype(np.float32)
    pp = satmc.make_poly_pairs(list(rv(pr["rw"], pr["rh"])), list(rv(pr["ow"], pr["oh"])), pr["rx"], pr["ry"], pr["rtheta"], pr["sd_x"], pr["sd_y"], pr["sd_theta"])
    d_pairs = torch.from_numpy(np.ascontiguousarray(pp).view(np.uint8).view(np.float32)).cuda()
    d_hits = torch.zeros(pp.size, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.count_fused_polygons(d_pairs, pp.size, 10_000, 7, d_hits)
elif mode == "cfg4":                                   # one pair, long items: k_count<DirectSrc, false, DEFER, MULTI>
    one = put(wl.cfg2_pair()); d_hits = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.count_fused(one, 1, 2_000_000_000, 7, d_hits)
elif mode == "cfg2":                                   # one pair x 1e6: k_count<DirectSrc, false, false, MULTI> (packed arrival counters)
    one = put(wl.cfg2_pair()); d_hits = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(6):
        ctx.count_fused(one, 1, 1_000_000, 7, d_hits)
elif mode == "adaptive":                               # the whole adaptive loop: k_count<IndirectSrc,...> + k_ztest_compact per iteration
    pairs = wl.dataset_pairs(100_000, 3)
    rb, poses, sds, pi, si, pos = wl.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-4, 1e-3, 1e-2], np.float32)
    dd = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rb, poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
    d_hits = torch.zeros(pairs.size, device="cuda")
    it, drawn = ctx.adaptive_run(dd[0], dd[1], pairs.size, dd[2], pairs.size, dd[3], dd[4], dd[5], pairs.size, dd[6], dd[7], 4,
                                 1_020_000, 1000, 20000, 100000, 7, d_hits)
    print("iterations", it, "samples", drawn)
elif mode in ("fused", "fused5", "ztest"):
    pairs = wl.dataset_pairs(100_000, 3, shape_variance=(mode == "fused5"))
    n = 1000 if mode == "ztest" else 10_000
    d_pairs = put(pairs); d_hits = torch.zeros(pairs.size, dtype=torch.int64, device="cuda")
    for _ in range(4):
        ctx.count_fused(d_pairs, pairs.size, n, 7, d_hits)
else:
    ndof = 3 if mode == "streamed3" else 5
    npairs, n = 16384, 32768
    z = torch.randn(ndof * npairs * n, device="cuda")
    pairs = wl.dataset_pairs(npairs, 9); d_pairs = put(pairs); d_hits = torch.zeros(npairs, dtype=torch.int64, device="cuda")
    for _ in range(4):
        ctx.count_streamed(d_pairs, npairs, z, npairs * n, ndof, n, d_hits, z_pair_stride=n)
torch.cuda.synchronize()
print(mode, "sum", float(d_hits.sum().item()), "launches", ctx.launch_count)
