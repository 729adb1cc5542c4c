#!/usr/bin/env python
"""Development probe (GPU box): throughput of the counting kernels on the BASELINE shapes + screening stats."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
if os.environ.get("SATMC_LIB"):                 # development only: compare builds of the same source
    satmc.LIB_PATH = os.environ["SATMC_LIB"]
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")


def put(a):
    a = np.ascontiguousarray(a)
    if a.dtype.fields is not None:
        a = a.view(np.float32)
    return torch.from_numpy(a).cuda()


def time_call(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
    print(torch.cuda.get_device_name(0))
    flags = [0] + ([satmc.SATMC_EXACT_ONLY] if "--exact" in sys.argv else [])
    if "--only" in sys.argv:
        if "--streamed" in sys.argv:
            streamed_part(ctx)
        return rest_part(ctx)
    cases = (("cfg3 1e5x1e4", wl.dataset_pairs(100_000, 3), 10_000),
                           ("cfg3 5dof 1e5x1e4", wl.dataset_pairs(100_000, 3, shape_variance=True), 10_000),
                           ("cfg5 slice 64000x1e5", wl.variance_sweep_pairs(1000, 5), 100_000),
                           ("cfg2 1x1e6", wl.cfg2_pair(), 1_000_000),
                           ("cfg4 slice 1x1e10", wl.cfg2_pair(), 10_000_000_000),
                           ("ztest 1e5x1000", wl.dataset_pairs(100_000, 3), 1000))
    if "--short" in sys.argv:
        cases = cases[:2] + cases[4:]
    for name, pairs, n in cases:
        d_pairs = put(pairs); d_hits = torch.zeros(pairs.size, dtype=torch.int64, device="cuda")
        for fl in flags:
            if fl and pairs.size * n > 2e10:
                continue
            ctx.exact_evals(reset=True)
            best, med = time_call(lambda: ctx.count_fused(d_pairs, pairs.size, n, 7, d_hits, flags=fl), reps=3 if pairs.size * n > 5e9 else 7)
            ev = ctx.exact_evals()
            tests = pairs.size * n
            print(f"fused  {name:24s} flags={fl} best {best:9.3f} ms  med {med:9.3f} ms  {tests / best / 1e6:10.2f} Gtests/s "
                  f" exact-eval frac {ev / (tests * (3 + 2 + 7 if pairs.size * n <= 5e9 else 2 + 3)):.2e}  p={d_hits.sum().item() / tests:.4f}")
    if "--short" not in sys.argv or "--streamed" in sys.argv:
        streamed_part(ctx)
    rest_part(ctx)


def streamed_part(ctx):
    # streamed: shared L2-resident bank (cfg5) and HBM-bound private slices
    pairs = wl.variance_sweep_pairs(1000, 5)
    d_pairs = put(pairs); d_hits = torch.zeros(pairs.size, dtype=torch.int64, device="cuda")
    for ndof in (3, 5):
        n = 100_000
        z = torch.randn(ndof * n, device="cuda")
        best, med = time_call(lambda: ctx.count_streamed(d_pairs, pairs.size, z, n, ndof, n, d_hits))
        print(f"streamed shared-bank ndof={ndof} {pairs.size}x{n}: best {best:.3f} ms {pairs.size * n / best / 1e6:.2f} Gtests/s")
    for ndof in (3, 5):
        npairs, n = int(os.environ.get("QB_NPAIRS", 16384)), int(os.environ.get("QB_N", 32768))
        torch.manual_seed(ndof)
        z = torch.randn(ndof * npairs * n, device="cuda")          # 6.4 / 10.7 GB
        pp = wl.dataset_pairs(npairs, 9); d_pp = put(pp); d_h = torch.zeros(npairs, dtype=torch.int64, device="cuda")
        best, med = time_call(lambda: ctx.count_streamed(d_pp, npairs, z, npairs * n, ndof, n, d_h, z_pair_stride=n))
        gb = ndof * 4 * npairs * n / 1e9
        print(f"streamed private ndof={ndof} {npairs}x{n}: best {best:.3f} ms {npairs * n / best / 1e6:.2f} Gtests/s  {gb / best * 1e3:.1f} GB/s"
              f"  hits {d_h.sum().item()}")
        del z


def rest_part(ctx):
    if "--sweep" in sys.argv:
        base = wl.dataset_pairs(10_000, 5)
        grid = np.array([0.01, 0.05, 0.15, 0.3]); vx, vy, vt = np.meshgrid(grid, grid, grid, indexing="ij")
        sig = np.sqrt(np.stack([vx.ravel(), vy.ravel(), vt.ravel()], 1)).astype(np.float32)
        d_b = put(base); d_s = sig; d_h = torch.zeros(base.size * 64, dtype=torch.int64, device="cuda")
        ctx.exact_evals(reset=True)
        best, med = time_call(lambda: ctx.count_fused_sweep(d_b, base.size, d_s, 64, 100_000, 7, d_h), reps=3)
        print(f"fused sweep cfg5 1e4 pairs x 64 cov x 1e5: best {best:.3f} ms {base.size * 64 * 1e5 / best / 1e6:.2f} Gtests/s "
              f"exact-eval frac {ctx.exact_evals() / (base.size * 64 * 1e5 * 5):.2e} p={d_h.sum().item() / (base.size * 64 * 1e5):.4f}")
    if "--poly" in sys.argv:
        rng = np.random.default_rng(0)
        base = wl.dataset_pairs(100_000, 3)
        def rect(w, h): return np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]], np.float32)
        def reg(k, r): a = 2 * np.pi * np.arange(k) / k; return np.stack([r * np.cos(a), r * np.sin(a)], 1).astype(np.float32)
        for name, robots, obstacles in (("4x4 (rectangles)", [rect(4.07, 1.74)] * base.size, [rect(p["ow"], p["oh"]) for p in base]),
                                        ("8x8 (octagons)", [reg(8, 2.0)] * base.size, [reg(8, 0.3 * (p["ow"] + p["oh"])) for p in base])):
            pp = satmc.make_poly_pairs(robots, obstacles, base["rx"], base["ry"], base["rtheta"], base["sd_x"], base["sd_y"], base["sd_theta"])
            d_pp = torch.from_numpy(pp.view(np.uint8).view(np.float32)).cuda(); d_h = torch.zeros(pp.size, dtype=torch.int64, device="cuda")
            best, med = time_call(lambda: ctx.count_fused_polygons(d_pp, pp.size, 10_000, 7, d_h), reps=3)
            print(f"fused polygons {name:18s} 1e5x1e4: best {best:.3f} ms {pp.size * 1e4 / best / 1e6:.2f} Gtests/s p={d_h.sum().item() / (pp.size * 1e4):.4f}")
    if "--ref" in sys.argv:
        from oracle.binding import RefGpu
        ref = RefGpu()
        pairs = wl.dataset_pairs(100_000, 3)
        rb, poses, sds, pi, si, pos = wl.reference_tables(pairs)
        ms, _ = ref.mc_time(rb, poses, sds, pi, si, pos, 1000, 10, 1)
        print(f"reference kernel cfg3 1e5 x (10 x 1000): {ms:.2f} ms  {1e9 / ms / 1e6:.3f} Gtests/s")


if __name__ == "__main__":
    main()
