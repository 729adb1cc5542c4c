"""CPU oracle checks (no GPU): analytic known answers, Philox KATs, SAT.py, golden vectors.

The reference ships no tests or golden vectors; the vectors under
tests/golden/ were produced by the UNMODIFIED reference compiled for sm_100a and run on a B200
(tools/make_golden.py), so replaying them here pins the C restatement to the reference binary.
"""
import glob
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rect(cx, cy, w, h, th=0.0):
    bx = np.array([-w / 2, w / 2, w / 2, -w / 2]); by = np.array([-h / 2, -h / 2, h / 2, h / 2])
    c, s = math.cos(th), math.sin(th)
    return np.stack([c * bx - s * by + cx, s * bx + c * by + cy], 1).reshape(8).astype(np.float32)


# ---- analytic cases -----------------------------------------------------
def test_identical_rectangles_collide(oracle):
    r = rect(0.3, -0.2, 2, 1, 0.4)
    assert oracle.convex_collide(r, r) == 1


def test_disjoint_rectangles_do_not_collide(oracle):
    assert oracle.convex_collide(rect(0, 0, 2, 1), rect(10, 0, 2, 1)) == 0
    assert oracle.convex_collide(rect(0, 0, 2, 1, 0.3), rect(0, 7, 2, 1, 1.1)) == 0


def test_edge_touching_counts_as_collision(oracle):
    # strict < at utils.cu:178: touching rectangles are NOT separated
    assert oracle.convex_collide(rect(0, 0, 2, 2), rect(2, 0, 2, 2)) == 1
    assert oracle.convex_collide(rect(0, 0, 2, 2), rect(2.0000005, 0, 2, 2)) == 0


def test_nan_counts_as_collision(oracle):
    r = rect(0, 0, 2, 2); q = rect(10, 0, 2, 2); q[0] = np.nan
    # every comparison involving the NaN projection is false -> never "separated" on those axes
    assert oracle.convex_collide(r, np.full(8, np.nan, np.float32)) == 1


def test_create_rect_order(oracle):
    np.testing.assert_array_equal(oracle.create_rect(4.0, 2.0), np.array([-2, -1, 2, -1, 2, 1, -2, 1], np.float32))


def test_cuda_trig_restatement_close_to_libm(oracle):
    xs = np.concatenate([np.random.default_rng(0).normal(0, 4, 5000), [0.0, 1e-30, 105614.9, 105615.0, 1e7, -3e9, 3e38]])
    for x in xs.astype(np.float32):
        assert abs(oracle.cuda_sinf(x) - math.sin(float(x))) < 2.5e-7
        assert abs(oracle.cuda_cosf(x) - math.cos(float(x))) < 2.5e-7
    assert math.isnan(oracle.cuda_sinf(float("inf"))) and math.isnan(oracle.cuda_cosf(float("nan")))


def test_sigma_zero_gives_zero_or_one(oracle, satmc):
    z = np.random.default_rng(1).standard_normal((3, 500)).astype(np.float32)
    p_hit = satmc.pairs_from_columns(1.0, 0.5, 0.3, 2.0, 1.0, 0, 0, 0)
    p_miss = satmc.pairs_from_columns(9.0, 0.5, 0.3, 2.0, 1.0, 0, 0, 0)
    assert oracle.count_streamed(p_hit, z) == 500
    assert oracle.count_streamed(p_miss, z) == 0


def test_axis_aligned_closed_form(oracle, satmc):
    # only sd_x != 0, theta = 0, |y0| < (h_r + h_o)/2  ->  p = Phi((a - x0)/s) - Phi((-a - x0)/s), a = (w_r + w_o)/2
    from math import erf, sqrt
    x0, s, wr, wo = 3.4, 0.8, 4.07, 2.3
    pair = satmc.pairs_from_columns(x0, 0.2, 0.0, wo, 1.1, s, 0.0, 0.0)
    n = 200_000
    z = np.random.default_rng(7).standard_normal((3, n)).astype(np.float32)
    a = (wr + wo) / 2
    Phi = lambda t: 0.5 * (1 + erf(t / sqrt(2)))
    p = Phi((a - x0) / s) - Phi((-a - x0) / s)
    k = oracle.count_streamed(pair, z)
    assert abs(k - n * p) < 4.9 * math.sqrt(n * p * (1 - p))


def test_three_and_five_dof_agree_when_shape_sigma_is_zero(oracle, workloads):
    pairs = workloads.dataset_pairs(20, seed=11)
    z5 = workloads.normal_bank(3000, 5, seed=12)
    a = oracle.count_streamed_batch(pairs, z5, 3000)
    b = oracle.count_streamed_batch(pairs, z5[:3].copy(), 3000)
    np.testing.assert_array_equal(a, b)


# ---- stop rule --------------------------------------------------------------------------------
def test_calc_slack_and_bins(oracle):
    assert oracle.calc_slack(1000, 0) == pytest.approx(math.log(40.0) / 1000, rel=1e-6)
    assert oracle.calc_slack(1000, 1000) == pytest.approx(math.log(40.0) / 1000, rel=1e-6)
    k, n = 300, 1000
    assert oracle.calc_slack(n, k) == pytest.approx(1.96 / n * math.sqrt(k - k * k / n), rel=1e-5)
    bins = [0.0, 0.01, 0.1, 1.0]
    assert oracle.get_bin(0.005, bins) == 0
    assert oracle.get_bin(0.05, bins) == 1
    assert oracle.get_bin(0.5, bins) == 2
    assert oracle.get_bin(0.01, bins) == 1          # a value on an edge goes to the higher bin


# ---- Philox4x32-10 known answers ------------------------------------------
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,out", KAT)
def test_philox_kat(oracle, ctr, key, out):
    np.testing.assert_array_equal(oracle.philox(ctr, key), np.array(out, np.uint32))


def test_fused_normals_are_standard_normal(oracle):
    for ndof in (3, 5):
        z = np.array([oracle.fused_normals(123, 7, i, ndof) for i in range(20000)])
        assert z.shape == (20000, ndof)
        assert np.all(np.abs(z.mean(0)) < 4.5 / math.sqrt(20000))
        assert np.all(np.abs(z.std(0) - 1) < 0.03)
        assert np.all(np.abs(np.corrcoef(z.T) - np.eye(ndof)) < 0.04)
        # consecutive samples are independent too (they share Philox blocks)
        assert abs(np.corrcoef(z[:-1, 0], z[1:, ndof - 1])[0, 1]) < 0.04


# ---- SAT.py (BASELINE config 1) ---------------------------------------------------------------------
def _sat_float64(r1, r2):
    """Independent float64 polygon-overlap check with true edge normals."""
    r1 = r1.reshape(-1, 4, 2).astype(np.float64); r2 = r2.reshape(-1, 4, 2).astype(np.float64)
    ok = np.ones(len(r1), bool); margin = np.full(len(r1), np.inf)
    for r in (r1, r2):
        for i in range(4):
            e = r[:, (i + 1) % 4] - r[:, i]
            nrm = np.stack([-e[:, 1], e[:, 0]], 1) / np.linalg.norm(e, axis=1, keepdims=True)
            p1 = np.einsum("nk,nck->nc", nrm, r1); p2 = np.einsum("nk,nck->nc", nrm, r2)
            gap = np.maximum(p2.min(1) - p1.max(1), p1.min(1) - p2.max(1))
            ok &= gap <= 0; margin = np.minimum(margin, np.abs(gap))
    return ok.astype(np.uint8), margin


def test_config1_sat_py_matches_c_and_float64(oracle, workloads):
    from oracle import SAT
    r1, r2 = workloads.cfg1_rect_pairs(10_000, seed=1)
    c = oracle.sat_batch(r1, r2)
    np.testing.assert_array_equal(c, SAT.collide_numpy(r1, r2))
    assert SAT.collide(r1[0], r2[0]) == c[0]
    ref64, margin = _sat_float64(r1, r2)
    away = margin > 1e-4
    np.testing.assert_array_equal(c[away], ref64[away])
    assert 0.05 < c.mean() < 0.6                    # a healthy mix of hits and misses


# ---- golden vectors produced by the compiled reference on a B200 -------------------------------------
def _golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.fail(f"golden fixture {name} missing (generate with tools/make_golden.py on the GPU box)")
    return np.load(path)


def test_golden_convex_collide(oracle):
    g = _golden("ref_convex_collide.npz")
    np.testing.assert_array_equal(oracle.sat_batch(g["r1"], g["r2"]), g["collide"].astype(np.uint8))


def test_golden_device_trig(oracle):
    g = _golden("ref_device_trig.npz")
    s = np.array([oracle.cuda_sinf(x) for x in g["x"]], np.float32)
    c = np.array([oracle.cuda_cosf(x) for x in g["x"]], np.float32)
    np.testing.assert_array_equal(s.view(np.uint32), g["sin"].view(np.uint32))
    np.testing.assert_array_equal(c.view(np.uint32), g["cos"].view(np.uint32))


def test_golden_rot_trans(oracle):
    g = _golden("ref_rot_trans.npz")
    out = np.stack([oracle.rot_trans(r, dx, dy, dt) for r, dx, dy, dt in zip(g["r_in"], g["dx"], g["dy"], g["dt"])])
    np.testing.assert_array_equal(out.view(np.uint32), g["r_out"].view(np.uint32))


def test_golden_sample_rectangle(oracle):
    g = _golden("ref_sample_rectangle.npz")
    n_per = int(g["n_per"])
    z = g["z"]
    for idx in range(z.shape[1]):
        i = idx // n_per
        out = oracle.sample_rectangle(g["r_in"][i], g["sd"][i], z[:, idx])
        np.testing.assert_array_equal(out.view(np.uint32), g["corners"][idx].view(np.uint32))


def test_golden_mc_kernel_counts(oracle):
    """Hit counts and done flags of the reference's own kernel on the normals it drew (recorded)."""
    g = _golden("ref_mc_kernel.npz")
    n_batch, n_samples = int(g["n_batch"]), int(g["n_samples"])
    z = g["z"]
    for gidx in range(g["positions"].shape[0]):
        pi, si = int(g["pose_idxs"][gidx]), int(g["sd_idxs"][gidx])
        zz = np.ascontiguousarray(z[:, gidx * n_batch:(gidx + 1) * n_batch])
        k, done = oracle.mc_thread(g["robot_base"], g["poses"][pi], g["std_devs"][si], g["positions"][gidx],
                                   int(g["cps_in"][gidx]), zz, n_batch, n_samples, g["bins"], g["bin_acc"])
        assert k == int(g["cps_out"][gidx]), gidx
        assert done == int(g["done"][gidx]), gidx


# ---- general convex polygons: the oracle's SAT against a DIFFERENT algorithm in float64 -----------------------------------
def _convex_polys_intersect_f64(A, B):
    """Two convex polygons (counter-clockwise [k,2] float64) intersect iff a vertex of one lies inside or on the other, or two
    edges cross.  No separating axes involved: an independent check of the SAT decisions.  Returns (intersect, margin) with
    margin = the smallest |signed distance| met by the tests (a tie indicator)."""
    def inside(P, Q):                                   # for each vertex of P: min over Q's edges of the signed distance (>= 0: inside)
        e = np.roll(Q, -1, axis=0) - Q
        nrm = np.stack([e[:, 1], -e[:, 0]], 1) / np.maximum(np.hypot(e[:, 0], e[:, 1]), 1e-300)[:, None]     # outward normals
        d = -np.einsum("pqk,qk->pq", P[:, None, :] - Q[None, :, :], nrm)      # inward distance of every vertex to every edge line
        return d.min(axis=1)
    ia, ib = inside(A, B), inside(B, A)
    hit = (ia >= 0).any() or (ib >= 0).any()
    margin = min(np.abs(ia).min(), np.abs(ib).min())
    ea, eb = np.roll(A, -1, axis=0) - A, np.roll(B, -1, axis=0) - B
    for i in range(len(A)):                             # proper crossings of edge pairs
        for j in range(len(B)):
            r, s_ = ea[i], eb[j]
            den = r[0] * s_[1] - r[1] * s_[0]
            if abs(den) < 1e-300:
                continue
            w = B[j] - A[i]
            t = (w[0] * s_[1] - w[1] * s_[0]) / den
            u = (w[0] * r[1] - w[1] * r[0]) / den
            if 0 <= t <= 1 and 0 <= u <= 1:
                hit = True
            margin = min(margin, max(min(abs(t), abs(1 - t)) * np.hypot(*r), 0) if 0 <= u <= 1 else np.inf,
                         max(min(abs(u), abs(1 - u)) * np.hypot(*s_), 0) if 0 <= t <= 1 else np.inf)
    return hit, margin


def test_polygon_oracle_matches_float64_geometry(oracle, satmc):
    """The polygon path has no counterpart in the reference (README.md:3 only says the method extends), so its contract is
    this repo's own.  Independent pin: the oracle's per-sample SAT decisions (float32, edge normals) equal a float64
    intersection test built on a different principle (vertex containment + edge crossings), for every sample that is not
    within 1e-4 of a tie.  The GPU path is compared with the oracle bit for bit in tests/test_gpu_polygons.py."""
    rng = np.random.default_rng(77)

    def convex(k, r):
        a = np.sort(rng.uniform(0, 2 * np.pi, k))
        while np.diff(np.concatenate([a, [a[0] + 2 * np.pi]])).max() > 0.9 * np.pi:      # keep the origin well inside
            a = np.sort(rng.uniform(0, 2 * np.pi, k))
        return np.stack([r * np.cos(a), r * np.sin(a)], 1).astype(np.float32)
    checked = ties = hits = 0
    for trial in range(40):
        kr, ko = int(rng.integers(3, 9)), int(rng.integers(3, 9))
        rob, obs = convex(kr, rng.uniform(0.5, 2.5)), convex(ko, rng.uniform(0.3, 2.0))
        dist, ang, th = rng.uniform(0.5, 4.0), rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)
        sd = rng.uniform(0.05, 0.8, 3)
        pp = satmc.make_poly_pairs([rob], [obs], dist * np.cos(ang), dist * np.sin(ang), th, sd[0], sd[1], sd[2])
        n = 400
        z = rng.standard_normal((3, n)).astype(np.float32)
        _, dec = oracle.poly_count_streamed(pp[0], z, want_decisions=True)
        p = pp[0]
        c, s = np.cos(np.float64(p["rtheta"])), np.sin(np.float64(p["rtheta"]))
        R64 = rob.astype(np.float64)
        A = np.stack([c * R64[:, 0] - s * R64[:, 1] + np.float64(p["rx"]), s * R64[:, 0] + c * R64[:, 1] + np.float64(p["ry"])], 1)
        O64 = obs.astype(np.float64)
        for i in range(n):
            dt = np.float64(z[2, i]) * np.float64(p["sd_theta"])
            cc, ss = np.cos(dt), np.sin(dt)
            B = np.stack([cc * O64[:, 0] - ss * O64[:, 1] + np.float64(z[0, i]) * np.float64(p["sd_x"]),
                          ss * O64[:, 0] + cc * O64[:, 1] + np.float64(z[1, i]) * np.float64(p["sd_y"])], 1)
            hit, margin = _convex_polys_intersect_f64(A, B)
            if margin < 1e-4:
                ties += 1
                continue
            checked += 1
            hits += int(hit)
            assert int(dec[i]) == int(hit), (trial, i, margin)
    assert checked > 12_000 and 0.05 < hits / checked < 0.95 and ties < 0.05 * (checked + ties), (checked, hits, ties)
