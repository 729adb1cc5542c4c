"""GPU box: the C oracle against the UNMODIFIED reference compiled for sm_100a (oracle/_ref).

This is what pins the oracle: every function of the restatement is compared bit-for-bit with the
reference binary on shared inputs, including the normals the reference's own cuRAND state drew.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_device_trig_restatement_is_bit_exact(oracle, refgpu, torch_cuda):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 3, 200_000), rng.uniform(-1e5, 1e5, 50_000), 10.0 ** rng.uniform(-38, 38, 50_000),
                        -(10.0 ** rng.uniform(-38, 38, 20_000)), [0.0, -0.0, 105615.0, np.inf, -np.inf, np.nan]]).astype(np.float32)
    s, c = refgpu.dev_sincos(x)
    so = np.array([oracle.cuda_sinf(v) for v in x], np.float32)
    co = np.array([oracle.cuda_cosf(v) for v in x], np.float32)
    nan = np.isnan(s)
    assert np.array_equal(np.isnan(so), nan) and np.array_equal(np.isnan(co), np.isnan(c))
    np.testing.assert_array_equal(bits(so[~nan]), bits(s[~nan]))
    np.testing.assert_array_equal(bits(co[~nan]), bits(c[~nan]))


def test_convex_collide_matches_reference(oracle, refgpu, workloads, torch_cuda):
    r1, r2 = workloads.cfg1_rect_pairs(10_000, seed=1)
    np.testing.assert_array_equal(oracle.sat_batch(r1, r2), refgpu.convex_collide(r1, r2).astype(np.uint8))


def test_rot_trans_matches_reference(oracle, refgpu, workloads, torch_cuda):
    rng = np.random.default_rng(3)
    n = 5000
    r = workloads.cfg1_rect_pairs(n, seed=9)[0]
    dx, dy, dt = (rng.normal(0, 3, n).astype(np.float32) for _ in range(3))
    out = np.stack([oracle.rot_trans(a, b, c, d) for a, b, c, d in zip(r, dx, dy, dt)])
    np.testing.assert_array_equal(bits(out), bits(refgpu.rot_trans(r, dx, dy, dt)))


def test_sample_rectangle_matches_reference_on_its_own_normals(oracle, refgpu, torch_cuda):
    rng = np.random.default_rng(4)
    n, n_per = 128, 40
    w, h = rng.uniform(0.1, 5, n), rng.uniform(0.1, 5, n)
    rin = np.stack([-w / 2, -h / 2, w / 2, -h / 2, w / 2, h / 2, -w / 2, h / 2], 1).astype(np.float32)
    sd = np.sqrt(rng.uniform(0, 0.3, (n, 5))).astype(np.float32)
    sd[::2, 3:] = 0
    z, corners = refgpu.sample_record(rin, sd, n_per, seed=11)
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05
    out = np.stack([oracle.sample_rectangle(rin[i // n_per], sd[i // n_per], z[:, i]) for i in range(n * n_per)])
    np.testing.assert_array_equal(bits(out), bits(corners))


def test_mc_kernel_counts_match_reference(oracle, refgpu, workloads, torch_cuda):
    pairs = workloads.dataset_pairs(700, seed=21, shape_variance=True)      # > 1 block of 512, ragged
    pairs["sd_w"][::2] = 0; pairs["sd_h"][::2] = 0
    robot_base, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    n_batch, n_samples = 300, 900
    rng = np.random.default_rng(5)
    cps_in = rng.integers(0, 600, pairs.size).astype(np.float32)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([2e-3, 1e-2, 2.5e-2], np.float32)
    cps, done, z = refgpu.mc_run(robot_base, poses, sds, pi, si, pos, cps_in, bins, acc, n_samples, n_batch, seed=3)
    for g in range(pairs.size):
        zz = np.ascontiguousarray(z[:, g * n_batch:(g + 1) * n_batch])
        k, d = oracle.mc_thread(robot_base, poses[g], sds[g], pos[g], int(cps_in[g]), zz, n_batch, n_samples, bins, acc)
        assert k == int(cps[g]) and d == int(done[g]), g
    assert 0 < done.sum() < pairs.size


def test_write_collision_probability(refgpu, torch_cuda):
    c = np.array([0, 1, 250, 999, 1000], np.float32)
    np.testing.assert_array_equal(refgpu.write_cp(c, 1000), c / np.float32(1000))
