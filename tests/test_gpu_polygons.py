"""GPU box: general convex polygons against the oracle's restatement, the rectangle path and an
independent float64 check."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def regular(k, r, phase=0.0):
    a = phase + 2 * np.pi * np.arange(k) / k
    return np.stack([r * np.cos(a), r * np.sin(a)], 1).astype(np.float32)


def rect_poly(w, h):
    return np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]], np.float32)


def random_convex(rng, k, r):
    a = np.sort(rng.uniform(0, 2 * np.pi, k))
    rad = r * rng.uniform(0.6, 1.0, k)
    pts = np.stack([rad * np.cos(a), rad * np.sin(a)], 1)
    # convex hull of points on a star-shaped ring is not guaranteed convex: project on the circle to keep convexity
    return np.stack([r * np.cos(a), r * np.sin(a)], 1).astype(np.float32) if k > 3 else pts.astype(np.float32)


def poly_count(ctx, dev, pp, z=None, n=None, seed=None, **kw):
    d_pairs = dev.put(pp.view(np.uint8).view(np.float32))
    d_hits = dev.zeros(pp.size, np.uint64)
    if z is not None:
        ctx.count_streamed_polygons(d_pairs, pp.size, dev.put(z.ravel()), z.shape[1], z.shape[1], d_hits, **kw)
    else:
        ctx.count_fused_polygons(d_pairs, pp.size, n, seed, d_hits, **kw)
    ctx.synchronize()
    return dev.get(d_hits, np.uint64)


def test_polygon_decisions_match_oracle_bit_for_bit(ctx, dev, oracle, satmc):
    rng = np.random.default_rng(5)
    robots, obstacles = [], []
    for i in range(60):
        kr, ko = rng.integers(3, 9), rng.integers(3, 9)
        robots.append(random_convex(rng, kr, rng.uniform(0.5, 2.5)))
        obstacles.append(random_convex(rng, ko, rng.uniform(0.3, 2.5)))
    for k in (3, 4, 6, 8, 4, 8):                                                   # equal counts: the straight-line variants
        robots.append(random_convex(rng, k, rng.uniform(0.5, 2.5)))
        obstacles.append(random_convex(rng, k, rng.uniform(0.3, 2.5)))
    robots += [regular(1, 0.0), regular(2, 1.0), rect_poly(4.07, 1.74)]          # point, segment, the reference robot
    obstacles += [regular(5, 1.0), regular(8, 1.5), regular(1, 0.0)]
    n = len(robots)
    d = rng.uniform(0.5, 5.0, n); ang = rng.uniform(0, 2 * np.pi, n)
    pp = satmc.make_poly_pairs(robots, obstacles, d * np.cos(ang), d * np.sin(ang), rng.uniform(0, 6.28, n),
                               rng.uniform(0.05, 0.6, n), rng.uniform(0.05, 0.6, n), rng.uniform(0.0, 0.6, n))
    for ns in (1, 33, 1000, 2051):
        z = rng.standard_normal((3, ns)).astype(np.float32)
        want = np.array([oracle.poly_count_streamed(pp[i], z) for i in range(n)], np.uint64)
        np.testing.assert_array_equal(poly_count(ctx, dev, pp, z=z), want)
    assert 0 < want.sum() < n * 2051


def test_polygon_path_agrees_with_rectangle_path_on_rectangles(ctx, dev, workloads, satmc):
    """4-vertex polygons built from the rectangle workload: true edge normals vs the reference's edge directions are the
    same axis set for rectangles, so counts agree except for samples within rounding of a tie."""
    pairs = workloads.dataset_pairs(400, seed=61)
    robots = [rect_poly(p["rw"], p["rh"]) for p in pairs]
    obstacles = [rect_poly(p["ow"], p["oh"]) for p in pairs]
    pp = satmc.make_poly_pairs(robots, obstacles, pairs["rx"], pairs["ry"], pairs["rtheta"], pairs["sd_x"], pairs["sd_y"], pairs["sd_theta"])
    n = 20_000
    z = workloads.normal_bank(n, 3, seed=62)
    k_poly = poly_count(ctx, dev, pp, z=z).astype(np.int64)
    d_hits = dev.zeros(pairs.size, np.uint64)
    ctx.count_streamed(dev.put(pairs), pairs.size, dev.put(z.ravel()), n, 3, n, d_hits)
    ctx.synchronize()
    k_rect = dev.get(d_hits, np.uint64).astype(np.int64)
    assert np.abs(k_poly - k_rect).max() <= 1 and (k_poly != k_rect).sum() <= 3
    # the fused polygon path consumes the same Philox stream as the fused rectangle path (3 normals per sample)
    kf_poly = poly_count(ctx, dev, pp, n=n, seed=9, sample_offset=5, pair_id_offset=3).astype(np.int64)
    ctx.count_fused(dev.put(pairs), pairs.size, n, 9, d_hits, sample_offset=5, pair_id_offset=3)
    ctx.synchronize()
    kf_rect = dev.get(d_hits, np.uint64).astype(np.int64)
    assert np.abs(kf_poly - kf_rect).max() <= 1 and (kf_poly != kf_rect).sum() <= 3


def test_polygon_fused_sharding_and_float64_probability(ctx, dev, satmc):
    rng = np.random.default_rng(8)
    robots = [regular(6, 1.5, 0.3)] * 4
    obstacles = [regular(5, 1.2, 0.1), regular(3, 1.0), regular(8, 0.8), rect_poly(2.0, 0.7)]
    pp = satmc.make_poly_pairs(robots, obstacles, [2.4, 2.0, 2.1, 2.3], [0.3, -0.5, 0.2, 0.0], [0.2, 1.0, 2.0, 0.5], 0.4, 0.3, 0.5)
    n, seed = 400_000, 17
    whole = poly_count(ctx, dev, pp, n=n, seed=seed)
    parts = poly_count(ctx, dev, pp, n=123_457, seed=seed) + poly_count(ctx, dev, pp, n=n - 123_457, seed=seed, sample_offset=123_457)
    np.testing.assert_array_equal(whole, parts)
    # independent float64 Monte Carlo with numpy (true normals, separating axis)
    def separated_on_normals_of(P, A, B):
        """[m] bool: some edge normal of polygon batch P separates A from B (float64)."""
        e = np.roll(P, -1, axis=1) - P
        nrm = np.stack([e[..., 1], -e[..., 0]], -1)
        pa = np.einsum("nik,njk->nij", nrm, A); pb = np.einsum("nik,njk->nij", nrm, B)
        return ((pa.max(2) < pb.min(2)) | (pb.max(2) < pa.min(2))).any(1)
    m = 200_000
    for i in range(pp.size):
        zz = rng.standard_normal((3, m))
        R = pp["robot"][i][:2 * pp["n_robot"][i]].reshape(-1, 2).astype(np.float64)
        O = pp["obstacle"][i][:2 * pp["n_obstacle"][i]].reshape(-1, 2).astype(np.float64)
        th = float(pp["rtheta"][i]); c0, s0 = math.cos(th), math.sin(th)
        Rw = np.stack([c0 * R[:, 0] - s0 * R[:, 1] + pp["rx"][i], s0 * R[:, 0] + c0 * R[:, 1] + pp["ry"][i]], 1)
        dt = zz[2] * pp["sd_theta"][i]; c, s = np.cos(dt)[:, None], np.sin(dt)[:, None]
        Ow = np.stack([c * O[None, :, 0] - s * O[None, :, 1] + (zz[0] * pp["sd_x"][i])[:, None],
                       s * O[None, :, 0] + c * O[None, :, 1] + (zz[1] * pp["sd_y"][i])[:, None]], 2)
        A = np.broadcast_to(Rw, (m,) + Rw.shape)
        hit = ~(separated_on_normals_of(A, A, Ow) | separated_on_normals_of(Ow, A, Ow))
        p_ref = hit.mean(); p = whole[i] / n
        pm = (p + p_ref) / 2
        assert abs(p - p_ref) < 4.9 * math.sqrt(pm * (1 - pm) * (1 / n + 1 / m)) + 1e-9, (i, p, p_ref)
    assert 0.01 < (whole / n).min() and (whole / n).max() < 0.99


def test_polygon_screening_equals_exact(ctx, dev, workloads, satmc):
    """The polygon screening pass (robot normals against the obstacle's bounding circle; inscribed circles) may only decide
    what the exact SAT would decide the same way.  Differential test, counts with screening == counts with every sample
    through the exact pass (SATMC_EXACT_ONLY), over: rectangles of the dataset prior, random convex polygons of mixed
    vertex counts placed around first contact, the same polygons given CLOCKWISE, non-convex stars, shapes whose local
    origin lies outside (inradius 0), points and segments; fused and streamed entry points."""
    rng = np.random.default_rng(31)
    EXACT = 0x2

    def both(pp, n, seed):
        ctx.exact_evals(reset=True)
        fast = poly_count(ctx, dev, pp, n=n, seed=seed)
        evals = ctx.exact_evals(reset=True)
        exact = poly_count(ctx, dev, pp, n=n, seed=seed, flags=EXACT)
        np.testing.assert_array_equal(fast, exact)
        return fast, evals

    # rectangles of the dataset prior: most samples must be decided by the screening pass
    pairs = workloads.dataset_pairs(20_000, seed=71)
    pp = satmc.make_poly_pairs([rect_poly(p["rw"], p["rh"]) for p in pairs], [rect_poly(p["ow"], p["oh"]) for p in pairs],
                               pairs["rx"], pairs["ry"], pairs["rtheta"], pairs["sd_x"], pairs["sd_y"], pairs["sd_theta"])
    fast, evals = both(pp, 4_000, 5)
    assert 0 < fast.sum() and 0 < evals < 0.4 * pairs.size * 4_000

    # mixed shapes around first contact
    robots, obstacles = [], []
    for i in range(3_000):
        kr, ko = rng.integers(1, 9), rng.integers(1, 9)
        a = random_convex(rng, max(kr, 3), rng.uniform(0.3, 2.5))[:kr]
        b = random_convex(rng, max(ko, 3), rng.uniform(0.2, 2.5))[:ko]
        shift = rng.uniform(-1.5, 1.5, 2).astype(np.float32) if i % 3 == 0 else np.zeros(2, np.float32)   # local origin off-centre / outside
        kind = i % 5
        if kind == 1:
            a, b = a[::-1].copy(), b[::-1].copy()                                                      # clockwise input
        if kind == 2 and ko >= 5:
            b = (b * np.where(np.arange(ko) % 2 == 0, 1.0, 0.35)[:, None]).astype(np.float32)           # non-convex star
        robots.append(a); obstacles.append((b + shift).astype(np.float32))
    n = len(robots)
    d = rng.uniform(0.0, 5.0, n); ang = rng.uniform(0, 2 * np.pi, n)
    sig = 10.0 ** rng.uniform(-3, -0.2, (3, n))
    pp = satmc.make_poly_pairs(robots, obstacles, d * np.cos(ang), d * np.sin(ang), rng.uniform(0, 6.28, n), sig[0], sig[1], sig[2])
    fast, evals = both(pp, 6_000, 6)
    frac = fast / 6_000
    assert (frac == 0).any() and (frac == 1).any() and ((frac > 0.05) & (frac < 0.95)).sum() > 100
    assert 0 < evals < n * 6_000                                    # and the screening pass did decide something

    # streamed entry point, ragged size, hostile normals
    z = rng.standard_normal((3, 3_001)).astype(np.float32)
    z[:, 7] = [np.nan, 0, 0]; z[:, 8] = [0, np.inf, 0]; z[:, 9] = [9.0, -9.0, 1e30]; z[:, 10] = [0, 0, np.nan]
    np.testing.assert_array_equal(poly_count(ctx, dev, pp[:500], z=z), poly_count(ctx, dev, pp[:500], z=z, flags=EXACT))


def test_polygon_screening_margin_at_its_own_boundary(ctx, dev, satmc):
    """Worst case for the separating test of the screening pass: the obstacle's bounding circle tangent to a robot edge
    line, free rotation (so that some samples point a vertex straight at the edge and the true gap is only the margin),
    position noise of the order of the margin (so that samples fall on both sides of the screening threshold).  Where
    the screening pass says "separated" the exact pass must find a separating axis: counts with == counts without."""
    rng = np.random.default_rng(41)
    EXACT = 0x2
    robots, obstacles, px, py, th, sig = [], [], [], [], [], []
    for i in range(3_000):
        kr, ko = rng.integers(3, 9), rng.integers(3, 9)
        a = random_convex(rng, kr, rng.uniform(0.3, 2.5)).astype(np.float64)
        b = random_convex(rng, ko, rng.uniform(0.2, 2.5)).astype(np.float64)
        if i % 2:
            b = b + rng.uniform(-0.8, 0.8, 2)                                    # bounding circle not centred on the shape
        t = rng.uniform(0, 2 * np.pi); c, s = np.cos(t), np.sin(t)
        aw = np.stack([c * a[:, 0] - s * a[:, 1], s * a[:, 0] + c * a[:, 1]], 1)    # rotated robot, before translation
        e = rng.integers(0, kr); j = (e + 1) % kr
        ed = aw[j] - aw[e]; u = np.array([ed[1], -ed[0]]); u /= np.hypot(*u)       # outward unit normal of edge e
        rho = np.hypot(b[:, 0], b[:, 1]).max()
        lam = rng.uniform(0.2, 0.8)
        foot = aw[e] + lam * ed                                                   # point of the edge nearest to the circle centre
        p = -(foot + rho * u)                                                     # puts the edge line at distance rho from the origin
        robots.append(a.astype(np.float32)); obstacles.append(b.astype(np.float32))
        px.append(p[0]); py.append(p[1]); th.append(t)
        sig.append(3e-5 * (rho + np.abs(p).max()) * 10.0 ** rng.uniform(-1, 1))
    sig = np.array(sig)
    pp = satmc.make_poly_pairs(robots, obstacles, px, py, th, sig, sig, 3.0)
    n = 20_000
    ctx.exact_evals(reset=True)
    fast = poly_count(ctx, dev, pp, n=n, seed=3)
    evals = ctx.exact_evals(reset=True)
    exact = poly_count(ctx, dev, pp, n=n, seed=3, flags=EXACT)
    np.testing.assert_array_equal(fast, exact)
    assert 0.02 * pp.size * n < evals < 0.98 * pp.size * n             # samples on both sides of the screening threshold


def test_polygon_screening_margin_inscribed_circles(ctx, dev, satmc):
    """Worst case for the "overlapping" test of the screening pass: two regular polygons facing each other edge to
    edge, no rotation, inscribed circles tangent, position noise of the order of the margin: where the screening pass
    says "overlapping" the true overlap is only the margin deep, and the exact pass must still find no separating axis."""
    rng = np.random.default_rng(43)
    EXACT = 0x2
    robots, obstacles, px, py, sig = [], [], [], [], []
    for i in range(2_000):
        kr, ko = rng.choice([3, 4, 5, 6, 8], 2)
        Rr, Ro = rng.uniform(0.3, 2.5), rng.uniform(0.2, 2.5)
        ph = rng.uniform(0, 2 * np.pi)
        alpha = ph + np.pi / kr                                        # outward direction of robot edge 0
        a, b = Rr * np.cos(np.pi / kr), Ro * np.cos(np.pi / ko)         # inradii
        robots.append(regular(kr, Rr, ph)); obstacles.append(regular(ko, Ro, alpha + np.pi - np.pi / ko))
        px.append(-(a + b) * np.cos(alpha)); py.append(-(a + b) * np.sin(alpha))
        sig.append(3e-5 * (Rr + Ro + a + b) * 10.0 ** rng.uniform(-1, 1))
    sig = np.array(sig)
    pp = satmc.make_poly_pairs(robots, obstacles, px, py, 0.0, sig, sig, 0.0)
    n = 20_000
    ctx.exact_evals(reset=True)
    fast = poly_count(ctx, dev, pp, n=n, seed=4)
    evals = ctx.exact_evals(reset=True)
    exact = poly_count(ctx, dev, pp, n=n, seed=4, flags=EXACT)
    np.testing.assert_array_equal(fast, exact)
    frac = fast / n
    assert ((frac > 0.2) & (frac < 0.8)).mean() > 0.8                  # the pairs really sit at first contact
    assert 0.3 * pp.size * n < evals < 0.98 * pp.size * n              # some samples decided by the inscribed circles, most not
