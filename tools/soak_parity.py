#!/usr/bin/env python
"""Extended differential run (GPU box): counts with the screening passes == counts with every sample through the exact
arithmetic, over more seeds and populations than the test suite affords.  Prints one line per case; exit code 1 on any
difference.  usage: soak_parity.py [n_seeds]"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
from test_gpu_polygons import random_convex
EXACT = 0x2
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).view(np.float32)).cuda()
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bad = 0


def rect_case(name, pairs, n, seed):
    global bad
    d = put(pairs); h = [torch.zeros(pairs.size, dtype=torch.int64, device="cuda") for _ in range(2)]
    ctx.exact_evals(reset=True)
    ctx.count_fused(d, pairs.size, n, seed, h[0]); ctx.synchronize(); ev = ctx.exact_evals(reset=True)
    ctx.count_fused(d, pairs.size, n, seed, h[1], flags=EXACT); ctx.synchronize()
    diff = int((h[0] != h[1]).sum().item()); bad += diff
    print(f"rect {name:28s} seed {seed:3d}: {pairs.size * n:.2e} tests, undecided {ev / (pairs.size * n):.2e}, mean p {h[0].sum().item() / (pairs.size * n):.4f}, differing pairs {diff}", flush=True)


def poly_case(name, pp, n, seed):
    global bad
    d = put(pp); h = [torch.zeros(pp.size, dtype=torch.int64, device="cuda") for _ in range(2)]
    ctx.exact_evals(reset=True)
    ctx.count_fused_polygons(d, pp.size, n, seed, h[0]); ctx.synchronize(); ev = ctx.exact_evals(reset=True)
    ctx.count_fused_polygons(d, pp.size, n, seed, h[1], flags=EXACT); ctx.synchronize()
    diff = int((h[0] != h[1]).sum().item()); bad += diff
    print(f"poly {name:28s} seed {seed:3d}: {pp.size * n:.2e} tests, to exact pass {ev / (pp.size * n):.3f}, mean p {h[0].sum().item() / (pp.size * n):.4f}, differing pairs {diff}", flush=True)


def sweep_case(name, pairs, sig, n, seed):
    global bad
    d = put(pairs); h = [torch.zeros(pairs.size * sig.shape[0], dtype=torch.int64, device="cuda") for _ in range(2)]
    ctx.exact_evals(reset=True)
    ctx.count_fused_sweep(d, pairs.size, sig, sig.shape[0], n, seed, h[0]); ctx.synchronize(); ev = ctx.exact_evals(reset=True)
    ctx.count_fused_sweep(d, pairs.size, sig, sig.shape[0], n, seed, h[1], flags=EXACT); ctx.synchronize()
    diff = int((h[0] != h[1]).sum().item()); bad += diff
    tests = pairs.size * sig.shape[0] * n
    print(f"sweep {name:27s} seed {seed:3d}: {tests:.2e} tests, undecided {ev / tests:.2e}, mean p {h[0].sum().item() / tests:.4f}, differing rows {diff}", flush=True)


def streamed_case(name, pairs, ndof, n, seed):
    global bad
    d = put(pairs); h = [torch.zeros(pairs.size, dtype=torch.int64, device="cuda") for _ in range(2)]
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    z = torch.randn(ndof * n, device="cuda", generator=g)
    ctx.count_streamed(d, pairs.size, z, n, ndof, n, h[0]); ctx.synchronize()
    ctx.count_streamed(d, pairs.size, z, n, ndof, n, h[1], flags=EXACT); ctx.synchronize()
    diff = int((h[0] != h[1]).sum().item()); bad += diff
    print(f"streamed {name:24s} seed {seed:3d}: {pairs.size * n:.2e} tests (shared bank, ndof {ndof}), mean p {h[0].sum().item() / (pairs.size * n):.4f}, differing pairs {diff}", flush=True)


def streamed_private_case(name, pairs, ndof, n, seed):
    """private HBM-resident slices: the bulk-tensor ring runs across short work items"""
    global bad
    d = put(pairs); h = [torch.zeros(pairs.size, dtype=torch.int64, device="cuda") for _ in range(2)]
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    z = torch.randn(ndof * pairs.size * n, device="cuda", generator=g)
    ctx.count_streamed(d, pairs.size, z, pairs.size * n, ndof, n, h[0], z_pair_stride=n); ctx.synchronize()
    ctx.count_streamed(d, pairs.size, z, pairs.size * n, ndof, n, h[1], z_pair_stride=n, flags=EXACT); ctx.synchronize()
    diff = int((h[0] != h[1]).sum().item()); bad += diff
    print(f"streamed {name:24s} seed {seed:3d}: {pairs.size * n:.2e} tests (private slices, ndof {ndof}), mean p {h[0].sum().item() / (pairs.size * n):.4f}, differing pairs {diff}", flush=True)


t0 = time.time()
rng = np.random.default_rng(2026)
for s in range(n_seeds):
    rect_case("dataset prior", wl.dataset_pairs(100_000, seed=1000 + s), 100_000, 11 + s)
    rect_case("dataset prior, 5-DoF", wl.dataset_pairs(100_000, seed=2000 + s, shape_variance=True), 50_000, 21 + s)
    p3 = wl.dataset_pairs(100_000, seed=3000 + s)
    p3["ow"] = rng.uniform(0.01, 0.3, p3.size); p3["rx"] *= 0.75; p3["ry"] *= 0.75
    for k in ("sd_x", "sd_y", "sd_theta"): p3[k] *= 0.05
    rect_case("thin obstacles, tiny sigma", p3, 100_000, 31 + s)
    p4 = wl.dataset_pairs(100_000, seed=4000 + s, max_variance=4.0)
    p4["rx"] = rng.uniform(-30, 30, p4.size); p4["ry"] = rng.uniform(-30, 30, p4.size)
    rect_case("huge sigma, far positions", p4, 100_000, 41 + s)
    p5 = wl.dataset_pairs(100_000, seed=5000 + s)
    p5["rx"] *= 1000.0; p5["ry"] *= 1000.0; p5["sd_x"] *= 1000.0; p5["sd_y"] *= 1000.0; p5["ow"] *= 1000.0; p5["oh"] *= 1000.0
    p5["rw"] *= 1000.0; p5["rh"] *= 1000.0
    rect_case("everything x 1000", p5, 50_000, 51 + s)
    m = 20_000
    robots = [random_convex(rng, rng.integers(3, 9), rng.uniform(0.3, 2.5)) for _ in range(m)]
    obstacles = [(random_convex(rng, rng.integers(3, 9), rng.uniform(0.2, 2.5)) + (rng.uniform(-1, 1, 2) if i % 2 else 0)).astype(np.float32) for i in range(m)]
    d = rng.uniform(0.0, 6.0, m); ang = rng.uniform(0, 2 * np.pi, m); sg = 10.0 ** rng.uniform(-2.5, -0.1, (3, m))
    pp = satmc.make_poly_pairs(robots, obstacles, d * np.cos(ang), d * np.sin(ang), rng.uniform(0, 6.28, m), sg[0], sg[1], sg[2])
    poly_case("random convex, mixed counts", pp, 100_000, 61 + s)
    sig = np.sqrt(rng.uniform(0.0, 0.3, (64, 3))).astype(np.float32)
    sig[:, 2] = rng.choice(np.sqrt(np.array([0.01, 0.05, 0.15, 0.3], np.float32)), 64)
    sweep_case("64 settings, 4 theta levels", wl.dataset_pairs(2_000, seed=6000 + s), sig, 20_000, 71 + s)
    streamed_case("dataset prior", wl.dataset_pairs(20_000, seed=7000 + s), 3, 100_000, 81 + s)
    streamed_case("dataset prior, 5-DoF", wl.dataset_pairs(20_000, seed=7500 + s, shape_variance=True), 5, 50_000, 91 + s)
    streamed_private_case("dataset prior", wl.dataset_pairs(5_000, seed=8000 + s), 3, 40_004 + 4 * s, 101 + s)
    streamed_private_case("dataset prior, 5-DoF", wl.dataset_pairs(5_000, seed=8500 + s, shape_variance=True), 5, 20_004 + 4 * s, 111 + s)
print(f"total differing pairs {bad}; {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
