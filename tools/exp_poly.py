"""dev probe: polygon screening vs exact-only, which pairs differ"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
from test_gpu_polygons import regular, rect_poly, random_convex
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).view(np.float32)).cuda()
rng = np.random.default_rng(5)
robots, obstacles = [], []
for i in range(60):
    kr, ko = rng.integers(3, 9), rng.integers(3, 9)
    robots.append(random_convex(rng, kr, rng.uniform(0.5, 2.5))); obstacles.append(random_convex(rng, ko, rng.uniform(0.3, 2.5)))
for k in (3, 4, 6, 8, 4, 8):
    robots.append(random_convex(rng, k, rng.uniform(0.5, 2.5))); obstacles.append(random_convex(rng, k, rng.uniform(0.3, 2.5)))
robots += [regular(1, 0.0), regular(2, 1.0), rect_poly(4.07, 1.74)]
obstacles += [regular(5, 1.0), regular(8, 1.5), regular(1, 0.0)]
n = len(robots)
d = rng.uniform(0.5, 5.0, n); ang = rng.uniform(0, 2 * np.pi, n)
pp = satmc.make_poly_pairs(robots, obstacles, d * np.cos(ang), d * np.sin(ang), rng.uniform(0, 6.28, n),
                           rng.uniform(0.05, 0.6, n), rng.uniform(0.05, 0.6, n), rng.uniform(0.0, 0.6, n))
for ns in (1, 33, 1000, 2051):
    z = rng.standard_normal((3, ns)).astype(np.float32)
    dz = torch.from_numpy(z.ravel()).cuda()
    out = []
    for flags in (0, 2):
        h = torch.zeros(n, dtype=torch.int64, device="cuda")
        ctx.count_streamed_polygons(put(pp), n, dz, ns, ns, h, flags=flags)
        ctx.synchronize(); out.append(h.cpu().numpy())
    bad = np.nonzero(out[0] != out[1])[0]
    print("ns", ns, "differing pairs", bad.tolist())
    for i in bad[:6]:
        print("  pair", i, "nr", int(pp["n_robot"][i]), "no", int(pp["n_obstacle"][i]), "fast", out[0][i], "exact", out[1][i])

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from binding import Oracle
orc = Oracle()
rng = np.random.default_rng(5)
# regenerate the same stream as above up to z
robots, obstacles = [], []
for i in range(60):
    kr, ko = rng.integers(3, 9), rng.integers(3, 9)
    robots.append(random_convex(rng, kr, rng.uniform(0.5, 2.5))); obstacles.append(random_convex(rng, ko, rng.uniform(0.3, 2.5)))
for k in (3, 4, 6, 8, 4, 8):
    robots.append(random_convex(rng, k, rng.uniform(0.5, 2.5))); obstacles.append(random_convex(rng, k, rng.uniform(0.3, 2.5)))
robots += [regular(1, 0.0), regular(2, 1.0), rect_poly(4.07, 1.74)]
obstacles += [regular(5, 1.0), regular(8, 1.5), regular(1, 0.0)]
d = rng.uniform(0.5, 5.0, n); ang = rng.uniform(0, 2 * np.pi, n)
pp = satmc.make_poly_pairs(robots, obstacles, d * np.cos(ang), d * np.sin(ang), rng.uniform(0, 6.28, n),
                           rng.uniform(0.05, 0.6, n), rng.uniform(0.05, 0.6, n), rng.uniform(0.0, 0.6, n))
for ns in (1, 33, 1000, 2051):
    z = rng.standard_normal((3, ns)).astype(np.float32)
    want = np.array([orc.poly_count_streamed(pp[i], z) for i in range(n)], np.uint64)
    h = torch.zeros(n, dtype=torch.int64, device="cuda")
    ctx.count_streamed_polygons(put(pp), n, torch.from_numpy(z.ravel()).cuda(), ns, ns, h, flags=0)
    ctx.synchronize(); got = h.cpu().numpy()
    bad = np.nonzero(got != want.astype(np.int64))[0]
    print("vs oracle ns", ns, "bad", bad.tolist())
    for i in bad:
        print("  pair", i, "nr", int(pp["n_robot"][i]), "no", int(pp["n_obstacle"][i]), "gpu", got[i], "oracle", want[i], {k: pp[k][i] for k in ("rx", "ry", "rtheta", "sd_x", "sd_y", "sd_theta")})
        # which samples: one at a time
        diff = []
        for s_ in range(ns):
            zz = np.ascontiguousarray(z[:, s_:s_ + 1])
            hh = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.count_streamed_polygons(put(pp[i:i + 1]), 1, torch.from_numpy(zz.ravel()).cuda(), 1, 1, hh, flags=0)
            ctx.synchronize()
            o = orc.poly_count_streamed(pp[i], zz)
            if int(hh.item()) != int(o): diff.append((s_, int(hh.item()), int(o), zz.ravel().tolist()))
        print("   single-sample diffs:", diff[:5], len(diff))
