"""GPU box: the adaptive z-test scheduler, the dataset front-end and the three drop-in programs end to end."""
import math
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "convex-2d-gpu-collision-detection_b200", "host")


@pytest.fixture(scope="module")
def programs():
    subprocess.check_call(["make", "-s", "-C", HOST, "all"])
    return {n: os.path.join(HOST, n) for n in ("generate_dataset", "ztest", "compute_collision_probability")}


def slack_int64_safe(oracle, n, k):
    """calcSlack (utils.cu:186-196) with k*k formed exactly: identical to the reference (and the oracle) while
    k <= 46340; above that the reference's int32 product wraps, which the library does
    not replicate."""
    if k == n or k == 0 or k <= 46340:
        return oracle.calc_slack(n, k)
    f = np.float32
    return float(f(1.96) / f(n) * np.sqrt(f(k) - f(k * k) / f(n)))


def test_adaptive_run_matches_stepwise_emulation(ctx, dev, oracle, workloads):
    """satmc_adaptive_run == the reference loop semantics (ztest.cu:328-388) emulated pair by pair with
    satmc_count_fused + the oracle's stop rule; results in input order, independent of finishing order."""
    pairs = workloads.dataset_pairs(400, seed=23)
    rb, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-3, 4e-3, 3e-3], np.float32)
    small, switch, large, max_samples, seed = 1000, 4000, 10000, 60000, 99
    d = {k: dev.put(v) for k, v in dict(rb=rb, poses=poses.ravel(), sds=sds.ravel(), pi=pi, si=si, pos=pos.ravel(), bins=bins, acc=acc).items()}
    d_cp = dev.zeros(pairs.size, np.float32); d_ns = dev.zeros(pairs.size, np.int32)
    iters, drawn = ctx.adaptive_run(d["rb"], d["poses"], pairs.size, d["sds"], pairs.size, d["pi"], d["si"], d["pos"], pairs.size,
                                    d["bins"], d["acc"], 4, max_samples, small, switch, large, seed, d_cp, d_ns)
    ctx.synchronize()
    cp, ns = dev.get(d_cp), dev.get(d_ns)
    # emulation
    d_pairs = dev.put(pairs); d_hits = dev.zeros(pairs.size, np.uint64)
    total = np.zeros(pairs.size, np.int64); n_samples = 0
    alive = np.ones(pairs.size, bool); want_cp = np.zeros(pairs.size, np.float32); want_ns = np.zeros(pairs.size, np.int32)
    want_drawn = 0; want_iters = 0
    while alive.any() and n_samples < max_samples:
        nb = small if n_samples < switch else large
        ctx.count_fused(d_pairs, pairs.size, nb, seed, d_hits, sample_offset=n_samples)
        ctx.synchronize()
        n_samples += nb; want_iters += 1; want_drawn += int(alive.sum()) * nb
        total += np.where(alive, dev.get(d_hits, np.uint64).astype(np.int64), 0)
        for g in np.nonzero(alive)[0]:
            k = int(total[g])
            p = np.float32(k) / np.float32(n_samples)
            if slack_int64_safe(oracle, n_samples, k) <= acc[oracle.get_bin(p, bins)]:
                alive[g] = False; want_cp[g] = np.float32(k) / np.float32(n_samples); want_ns[g] = n_samples
    left = np.nonzero(alive)[0]
    want_cp[left] = total[left].astype(np.float32) / np.float32(n_samples); want_ns[left] = n_samples
    np.testing.assert_array_equal(ns, want_ns)
    np.testing.assert_array_equal(cp, want_cp)
    assert (iters, drawn) == (want_iters, want_drawn)
    assert len(set(want_ns.tolist())) > 3 and left.size > 0        # pairs really stop at different times, some never


def test_adaptive_run_vs_reference_loop(ctx, dev, refgpu, workloads):
    """The library's adaptive scheduler against the reference's own loop (its kernel + thrust::count +
    thrust::sort_by_key, oracle/ref_gpu.cu::ref_adaptive_batch): independent RNGs, so statistical agreement:
    |p1 - p2| <= 4.9 sqrt(2 p (1-p) / 1000) + 2e-3 (every pair used >= 1000 samples in both)."""
    pairs = workloads.dataset_pairs(3000, seed=29)
    rb, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-4, 1e-3, 1e-2], np.float32)
    max_samples = 120_000
    cp_ref, ms_ref, drawn_ref = refgpu.adaptive_batch(rb, poses, sds, pi, si, pos, bins, acc, max_samples, seed=4)
    d = {k: dev.put(v) for k, v in dict(rb=rb, poses=poses.ravel(), sds=sds.ravel(), pi=pi, si=si, pos=pos.ravel(), bins=bins, acc=acc).items()}
    d_cp = dev.zeros(pairs.size, np.float32)
    iters, drawn = ctx.adaptive_run(d["rb"], d["poses"], pairs.size, d["sds"], pairs.size, d["pi"], d["si"], d["pos"], pairs.size,
                                    d["bins"], d["acc"], 4, max_samples, 1000, 20000, 100000, 12345, d_cp)
    ctx.synchronize()
    cp = dev.get(d_cp)
    p = (cp + cp_ref) / 2
    assert np.all(np.abs(cp - cp_ref) <= 4.9 * np.sqrt(2 * p * (1 - p) / 1000) + 2e-3), np.abs(cp - cp_ref).max()
    assert 0.5 < drawn / drawn_ref < 2.0                 # both schedules draw a similar number of samples
    assert (cp_ref > 0).mean() > 0.02


def test_sample_positions_ring_prior(ctx, dev, workloads):
    rng = np.random.default_rng(0)
    n_poses, n_std, n = 50, 40, 200_000
    poses = np.stack([rng.uniform(0.1, 5, n_poses), rng.uniform(0.1, 5, n_poses), rng.uniform(0, 6.28, n_poses)], 1).astype(np.float32)
    sds = np.sqrt(rng.uniform(0, 0.3, (n_std, 5))).astype(np.float32)
    d_pos = dev.zeros(2 * n, np.float32); d_pi = dev.zeros(n, np.float32); d_si = dev.zeros(n, np.float32)
    r_off, spread = (4.07 + 1.74) / 4, 4.0
    ctx.sample_positions(dev.put(poses.ravel()), n_poses, dev.put(sds.ravel()), n_std, n, r_off, spread, 5, d_pos, d_pi, d_si)
    ctx.synchronize()
    pos = dev.get(d_pos).reshape(n, 2); pi = dev.get(d_pi).astype(int); si = dev.get(d_si).astype(int)
    assert pi.min() == 0 and pi.max() == n_poses - 1 and si.min() == 0 and si.max() == n_std - 1
    assert np.abs(np.bincount(pi, minlength=n_poses) / n - 1 / n_poses).max() < 5 * math.sqrt(1 / n_poses / n)
    # invert generate_dataset.cu:215-216: pos = (cos t * (A + shift), sin t * (B + shift)), shift ~ N(0, (sx+sy)/2 * spread)
    A = poses[pi, 0] / 2 + r_off + 2.35 + sds[si, 0]; B = poses[pi, 1] / 2 + r_off + 2.35 + sds[si, 1]
    sig = (sds[si, 0] + sds[si, 1]) / 2 * spread
    # E[x^2 / (A^2 + sig^2) ] = E[cos^2] = 1/2 (theta uniform, shift independent)
    ex = (pos[:, 0] ** 2 / (A ** 2 + sig ** 2)).mean(); ey = (pos[:, 1] ** 2 / (B ** 2 + sig ** 2)).mean()
    assert abs(ex - 0.5) < 0.01 and abs(ey - 0.5) < 0.01
    ang = np.arctan2(pos[:, 1] / B, pos[:, 0] / A)
    assert abs(np.cos(ang).mean()) < 0.01 and abs(np.sin(ang).mean()) < 0.01
    # deterministic in (seed, stream)
    d_pos2 = dev.zeros(2 * 1000, np.float32)
    ctx.sample_positions(dev.put(poses.ravel()), n_poses, dev.put(sds.ravel()), n_std, 1000, r_off, spread, 5, d_pos2, d_pi, d_si)
    ctx.synchronize()
    np.testing.assert_array_equal(dev.get(d_pos2).reshape(1000, 2), pos[:1000])


def load_data_like_balance_datasets(data_dir):
    """The loader of the reference's consumer (balance_datasets.py:6-13), numpy part only."""
    out = []
    for f in sorted(os.listdir(data_dir)):
        if f.endswith(".npy") and not f.startswith("poses") and not f.startswith("variance") and not f.startswith("checkpoint"):
            out.append(np.load(os.path.join(data_dir, f)))
    return np.concatenate(out, axis=0)


def test_three_programs_end_to_end(programs, tmp_path, ctx, dev, workloads):
    data = tmp_path / "data"
    B, nb = 3000, 2
    r = subprocess.run([programs["generate_dataset"], "--data_dir", str(data), "-n", str(nb), "-b", str(B), "--num_poses", "500",
                        "--num_variances", "400", "--max_samples", "40000", "--seed", "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    poses = np.load(data / "poses.npy"); var = np.load(data / "variances.npy")
    assert poses.shape == (500, 3) and var.shape == (400, 5) and poses.dtype == np.float32
    assert poses[:, :2].min() >= 0.1 and poses[:, :2].max() <= 5 and poses[:, 2].max() <= 6.2832
    assert var[:, :3].max() <= 0.3 and np.all(var[:, 3:] == 0)                      # no --shape_variance
    np.testing.assert_array_equal(np.load(data / "meta" / "accuracy_bins.npy"), np.array([0, 0.01, 0.1, 1], np.float32))
    np.testing.assert_array_equal(np.load(data / "meta" / "bin_accuracy.npy"), np.array([1e-4, 1e-3, 1e-2], np.float32))
    rows = load_data_like_balance_datasets(str(data))                               # consumer contract: [N,5] float32
    assert rows.shape == (nb * B, 5) and rows.dtype == np.float32
    cp = rows[:, 2]
    assert cp.min() >= 0 and cp.max() <= 1 and 0.02 < (cp > 0).mean() < 0.9
    assert np.all(rows[:, 3] == np.floor(rows[:, 3])) and rows[:, 3].max() < 400 and rows[:, 4].max() < 500
    sel = rows[(rows[:, 3] == rows[0, 3]) & (rows[:, 4] == rows[0, 4])]             # the notebook's selector (cell 0)
    assert sel.shape[0] >= 1
    # resume: --start_batch_count continues the numbering and does not touch earlier files
    before = np.load(data / "0.npy")
    r = subprocess.run([programs["generate_dataset"], "--data_dir", str(data), "-n", "1", "-b", str(B), "-s", "2", "--pose_dir",
                        str(data / "poses.npy"), "--variance_dir", str(data / "variances.npy"), "--max_samples", "40000",
                        "--seed", "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (data / "2.npy").exists()
    np.testing.assert_array_equal(np.load(data / "0.npy"), before)

    # ztest on the rows of batch 0: same tables, independent samples -> probabilities agree statistically
    b0 = np.load(data / "0.npy")
    (data / "tmp").mkdir()
    np.save(data / "tmp" / "0.npy", np.ascontiguousarray(b0[:, [0, 1, 3, 4]]))
    os.remove(data / "2.npy"); os.remove(data / "1.npy")
    out = tmp_path / "zt.npy"
    r = subprocess.run([programs["ztest"], "--data_dir", str(data), "--data_file_out", str(out), "--max_samples", "40000",
                        "--seed", "11", "--meta_dir", str(data / "meta")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    zt = np.load(out)
    assert zt.shape == b0.shape
    np.testing.assert_array_equal(zt[:, [0, 1, 3, 4]], b0[:, [0, 1, 3, 4]])          # input order, inputs echoed
    # both estimates used >= 1000 samples: |p1 - p2| <= 4.9 * sqrt(2 p (1-p) / 1000) + 2e-3
    p = (zt[:, 2] + b0[:, 2]) / 2
    assert np.all(np.abs(zt[:, 2] - b0[:, 2]) <= 4.9 * np.sqrt(2 * p * (1 - p) / 1000) + 2e-3)
    r = subprocess.run([programs["ztest"], "--data_dir", str(data), "--data_file_out", str(tmp_path / "cps.npy"), "--max_samples",
                        "40000", "--seed", "11", "--cps_only", "1", "--meta_dir", str(data / "meta")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    np.testing.assert_array_equal(np.load(tmp_path / "cps.npy"), zt[:, 2])           # same seed -> identical, 1-D

    # compute_collision_probability: data_in/<b>.npy -> data_out/<start+b>.npy, tables from data_out
    din = tmp_path / "din"; din.mkdir()
    np.save(din / "0.npy", np.ascontiguousarray(b0[:1000, [0, 1, 3, 4]]))
    np.save(din / "1.npy", np.ascontiguousarray(b0[1000:2500, [0, 1, 3, 4]]))
    r = subprocess.run([programs["compute_collision_probability"], "--data_in", str(din), "--data_out", str(data), "--max_samples",
                        "40000", "--shuffle", "0", "--seed", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    o1, o2 = np.load(data / "1.npy"), np.load(data / "2.npy")                        # data_out already held 0.npy -> start = 1
    assert o1.shape == (1000, 5) and o2.shape == (1500, 5)
    np.testing.assert_array_equal(o1[:, [0, 1, 3, 4]], b0[:1000, [0, 1, 3, 4]])
    pp = (o1[:, 2] + b0[:1000, 2]) / 2
    assert np.all(np.abs(o1[:, 2] - b0[:1000, 2]) <= 4.9 * np.sqrt(2 * pp * (1 - pp) / 1000) + 2e-3)


def test_generate_dataset_is_reproducible_and_gpu_count_invariant(programs, tmp_path, torch_cuda):
    """Same --seed -> identical files; with >= 2 GPUs, --gpus 2 writes the same files as --gpus 1."""
    common = ["-n", "3", "-b", "1500", "--num_poses", "200", "--num_variances", "200", "--max_samples", "30000", "--seed", "5"]
    outs = []
    runs = [["--gpus", "1"], ["--gpus", "1"]] + ([["--gpus", "2"]] if torch_cuda.cuda.device_count() >= 2 else [])
    for k, extra in enumerate(runs):
        d = tmp_path / f"run{k}"
        r = subprocess.run([programs["generate_dataset"], "--data_dir", str(d)] + common + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        outs.append([np.load(d / f"{b}.npy") for b in range(3)])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            np.testing.assert_array_equal(a, b)
    assert not np.array_equal(outs[0][0], outs[0][1])              # different batches differ


def test_ztest_and_compute_cp_gpu_count_invariant(programs, tmp_path, torch_cuda, workloads):
    """ztest / compute_collision_probability --gpus 2 write the same files as --gpus 1 (needs >= 2 GPUs)."""
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    data = tmp_path / "data"
    r = subprocess.run([programs["generate_dataset"], "--data_dir", str(data), "-n", "1", "-b", "2001", "--num_poses", "100",
                        "--num_variances", "100", "--max_samples", "30000", "--seed", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b0 = np.load(data / "0.npy")
    (data / "tmp").mkdir()
    np.save(data / "tmp" / "0.npy", np.ascontiguousarray(b0[:, [0, 1, 3, 4]]))
    outs = []
    for g in ("1", "2"):
        out = tmp_path / f"zt{g}.npy"
        r = subprocess.run([programs["ztest"], "--data_dir", str(data), "--data_file_out", str(out), "--max_samples", "30000", "--seed", "9",
                            "--meta_dir", str(data / "meta"), "--gpus", g], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        outs.append(np.load(out))
    np.testing.assert_array_equal(outs[0], outs[1])
    din = tmp_path / "din"; din.mkdir()
    np.save(din / "0.npy", np.ascontiguousarray(b0[:777, [0, 1, 3, 4]]))
    res = []
    for g in ("1", "2"):
        dout = tmp_path / f"dout{g}"; (dout / "meta").mkdir(parents=True)
        for f in ("poses.npy", "variances.npy"):
            np.save(dout / f, np.load(data / f))
        for f in ("accuracy_bins.npy", "bin_accuracy.npy"):
            np.save(dout / "meta" / f, np.load(data / "meta" / f))
        r = subprocess.run([programs["compute_collision_probability"], "--data_in", str(din), "--data_out", str(dout), "--max_samples",
                            "30000", "--seed", "4", "--gpus", g], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        res.append(np.load(dout / "0.npy"))
    np.testing.assert_array_equal(res[0], res[1])


def test_program_errors(programs, tmp_path):
    r = subprocess.run([programs["ztest"], "--data_dir", str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not exist" in r.stdout
    r = subprocess.run([programs["generate_dataset"], "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "unrecognised option" in r.stderr
    r = subprocess.run([programs["generate_dataset"], "--help"], capture_output=True, text=True)
    assert r.returncode == 1 and "--num_batches" in r.stdout


def test_plain_c_client(satmc, tmp_path):
    """tests/c/abi_example.c (C99, no torch, no C++) runs the fused path through the C ABI on the GPU."""
    libdir = os.path.dirname(satmc.LIB_PATH)
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", satmc.INCLUDE_DIR,
                           os.path.join(ROOT, "tests", "c", "abi_example.c"), "-o", exe, "-L", libdir, "-lsatmc", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "hits" in r.stdout
