"""GPU box: the multi-GPU split behind the C ABI (satmc_group_*), the single-launch counting scheme and the
reference kernel contract with an arbitrary robot quad.

A group of world size 1 runs everywhere; the tests that need two devices skip on a one-GPU box and are also
exercised by bench.py at N >= 2 (strong.cfg3 / strong.cfg4 / sharding_check / programs_gpu_count_invariance)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXACT = 0x2


def fused(ctx, dev, pairs, n, seed, sample_offset=0, pair_id_offset=0, flags=0):
    d_pairs = dev.put(pairs)
    d_hits = dev.zeros(pairs.size, np.uint64)
    ctx.count_fused(d_pairs, pairs.size, n, seed, d_hits, sample_offset=sample_offset, pair_id_offset=pair_id_offset, flags=flags)
    ctx.synchronize()
    return dev.get(d_hits, np.uint64)


def n_gpus(torch):
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_group_counts_equal_single_gpu(ctx, dev, satmc, workloads, torch_cuda, world):
    """Both shard modes, host buffers and resident inputs: every split returns the single-GPU counts bit for bit."""
    if n_gpus(torch_cuda) < world:
        pytest.skip(f"needs {world} GPUs")
    pairs = workloads.dataset_pairs(1003, seed=41, shape_variance=True)
    pairs["sd_w"][::2] = 0; pairs["sd_h"][::2] = 0
    n, seed, off, pid = 4099, 77, 123_456_789_011, 17
    want = fused(ctx, dev, pairs, n, seed, sample_offset=off, pair_id_offset=pid)
    with satmc.Group(devices=list(range(world))) as g:
        assert g.world == world and g.local_count == world and g.rank(world - 1) == world - 1
        for mode in (satmc.SHARD_BY_PAIR, satmc.SHARD_BY_SAMPLE_RANGE):
            got = g.count_fused_host(pairs, n, seed, mode, sample_offset=off, pair_id_offset=pid)
            np.testing.assert_array_equal(got, want, err_msg=f"host, mode {mode}")
        # sample ranges, few pairs, one process: the kernels finish into device 0's memory over NVLink (no collective);
        # with that path switched off the same call goes through ncclAllReduce -- same counts
        if world > 1:
            assert g.last_exchange() == ("peer_atomics" if world >= 6 else "nccl")        # automatic choice
            g.set_peer_reduce(True)
            got = g.count_fused_host(pairs, n, seed, satmc.SHARD_BY_SAMPLE_RANGE, sample_offset=off, pair_id_offset=pid)
            assert g.last_exchange() == "peer_atomics"
            np.testing.assert_array_equal(got, want, err_msg="host, sample ranges through peer atomics")
            g.set_peer_reduce(False)
            got = g.count_fused_host(pairs, n, seed, satmc.SHARD_BY_SAMPLE_RANGE, sample_offset=off, pair_id_offset=pid)
            assert g.last_exchange() == "nccl"
            np.testing.assert_array_equal(got, want, err_msg="host, sample ranges through NCCL")
            g.set_peer_reduce(None)
        else:
            assert g.last_exchange() == "none"
        # resident inputs: the full pair array on every device, capacity-sized counters
        cap = g.hits_capacity(pairs.size)
        dp, dh = [], []
        for l in range(world):
            with torch_cuda.cuda.device(l):
                dp.append(torch_cuda.from_numpy(np.ascontiguousarray(pairs).view(np.float32)).cuda())
                dh.append(torch_cuda.zeros(cap, dtype=torch_cuda.int64, device=f"cuda:{l}"))
        for mode in (satmc.SHARD_BY_PAIR, satmc.SHARD_BY_SAMPLE_RANGE):
            for h in dh:
                h.fill_(-1)
            g.count_fused(dp, pairs.size, n, seed, mode, dh, sample_offset=off, pair_id_offset=pid)
            g.synchronize()
            for l in range(world):
                np.testing.assert_array_equal(dh[l][:pairs.size].cpu().numpy().view(np.uint64), want, err_msg=f"device {l}, mode {mode}")
        g.set_timing(True)
        g.count_fused(dp, pairs.size, n, seed, satmc.SHARD_BY_SAMPLE_RANGE, dh)
        k_ms, c_ms = g.last_times()
        assert k_ms > 0 and c_ms >= 0
    assert want.sum() > 0


def test_group_single_pair_cfg4_slice(ctx, dev, satmc, workloads, torch_cuda):
    """cfg 4 in small: one pair, 2e9 samples by sample range over every GPU of the box == one GPU."""
    world = min(n_gpus(torch_cuda), 8)
    one = workloads.cfg2_pair()
    want = fused(ctx, dev, one, 2_000_000_000, 4)
    with satmc.Group(devices=list(range(world))) as g:
        g.set_peer_reduce(True)
        got = g.count_fused_host(one, 2_000_000_000, 4, satmc.SHARD_BY_SAMPLE_RANGE)
        assert g.last_exchange() == ("peer_atomics" if world > 1 else "none")
        g.set_peer_reduce(False)
        got2 = g.count_fused_host(one, 2_000_000_000, 4, satmc.SHARD_BY_SAMPLE_RANGE)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got2, want)
    assert abs(int(want[0]) / 2e9 - 0.166) < 0.01


def adaptive_single(ctx, dev, workloads, pairs, max_samples, seed, stream_offset, tight=False):
    rb, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32)
    acc = np.array([1e-5, 1e-5, 1e-5], np.float32) if tight else np.array([1e-3, 3e-3, 1e-2], np.float32)
    d = [dev.put(a) for a in (rb, poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
    d_cp = dev.zeros(pairs.size, np.float32)
    it, drawn = ctx.adaptive_run(d[0], d[1], pairs.size, d[2], pairs.size, d[3], d[4], d[5], pairs.size, d[6], d[7], 4, max_samples,
                                 1000, 20000, 100000, seed, d_cp, stream_id_offset=stream_offset)
    ctx.synchronize()
    return dev.get(d_cp), it, drawn, (rb, poses, sds, pi, si, pos, bins, acc)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_group_adaptive_rows_interleaved(ctx, dev, satmc, workloads, torch_cuda, world):
    """satmc_group_adaptive_run_host (rows dealt round-robin, devices in lockstep) == satmc_adaptive_run on one GPU:
    row i always draws Philox stream offset + i, whatever device it lands on."""
    if n_gpus(torch_cuda) < world:
        pytest.skip(f"needs {world} GPUs")
    pairs = workloads.dataset_pairs(5001, seed=43)
    want, it1, drawn1, (rb, poses, sds, pi, si, pos, bins, acc) = adaptive_single(ctx, dev, workloads, pairs, 230_000, 9, 1000)
    with satmc.Group(devices=list(range(world))) as g:
        g.set_tables(rb, poses, sds, bins, acc)
        cp, it, drawn = g.adaptive_run_host(pi, si, pos, 230_000, 1000, 20000, 100000, 9, stream_id_offset=1000)
    np.testing.assert_array_equal(cp, want)
    assert drawn == drawn1 and it == it1
    assert 0 < (want > 0).mean() < 1
    # the same with a budget that stops the loop before every row is done (the leftovers are finalised as they stand)
    want2, it2, drawn2, _ = adaptive_single(ctx, dev, workloads, pairs, 25_000, 9, 1000, tight=True)
    with satmc.Group(devices=list(range(world))) as g:
        g.set_tables(rb, poses, sds, bins, np.array([1e-5, 1e-5, 1e-5], np.float32))
        cp2, it_g, drawn_g = g.adaptive_run_host(pi, si, pos, 25_000, 1000, 20000, 100000, 9, stream_id_offset=1000)
    np.testing.assert_array_equal(cp2, want2)
    assert it_g == it2 and drawn_g == drawn2


def test_single_launch_counters_leave_no_state_behind(ctx, dev, oracle, workloads):
    """Calls with several work items per counter accumulate into a scratch array that must be all zero again after every
    launch (the last block moves the totals out): repeat, change sizes (the scratch grows), mix with SATMC_ACCUMULATE,
    poison the output first, and compare with the CPU restatement."""
    one = workloads.cfg2_pair()
    few = workloads.dataset_pairs(37, seed=51, shape_variance=True)
    want1 = oracle.count_fused_batch(one, 300_000, 5)
    want37 = oracle.count_fused_batch(few, 50_000, 6, sample_offset=3)
    l0 = ctx.launch_count
    fused(ctx, dev, one, 300_000, 5)
    assert ctx.launch_count - l0 == 1                                # one kernel, no memset node
    for rep in range(3):
        d_h = dev.zeros(1, np.uint64); d_h.fill_(-7)                 # garbage in the output must not matter
        ctx.count_fused(dev.put(one), 1, 300_000, 5, d_h); ctx.synchronize()
        np.testing.assert_array_equal(dev.get(d_h, np.uint64), want1)
        ctx.count_fused(dev.put(one), 1, 300_000, 5, d_h, flags=0x1); ctx.synchronize()         # accumulate on top
        np.testing.assert_array_equal(dev.get(d_h, np.uint64), 2 * want1)
        np.testing.assert_array_equal(fused(ctx, dev, few, 50_000, 6, sample_offset=3), want37)  # more counters: scratch regrows once
        z = workloads.normal_bank(70_001, 5, seed=52)
        d_hs = dev.zeros(few.size, np.uint64); d_hs.fill_(-1)
        ctx.count_streamed(dev.put(few), few.size, dev.put(z.ravel()), 70_001, 5, 70_001, d_hs); ctx.synchronize()
        np.testing.assert_array_equal(dev.get(d_hs, np.uint64), oracle.count_streamed_batch(few, z, 70_001))


def test_mc_step_honours_any_robot_quad(ctx, dev, oracle, workloads):
    """The reference kernel transforms whatever 8 floats robot_base holds (ztest.cu:148-149).  A quad that is not
    create_rect(w, h) -- here an off-centre, sheared footprint -- must be evaluated on its own corners (exact arithmetic
    for every sample); the same rectangle listed from another corner must reproduce the standard order's counts."""
    pairs = workloads.dataset_pairs(64, seed=61)
    rb, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.zeros(3, np.float32)
    n_batch, seed = 2048, 13

    def step(robot_base):
        d = [dev.put(a) for a in (np.asarray(robot_base, np.float32), poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
        d_cps = dev.zeros(pairs.size, np.float32); d_done = dev.zeros(pairs.size, np.int32)
        ctx.mc_step(d[0], d[1], pairs.size, d[2], pairs.size, d[3], d[4], d[5], d_cps, d[6], d[7], 4, d_done, 0, n_batch, n_batch,
                    pairs.size, seed, 5)
        ctx.synchronize()
        return dev.get(d_cps).astype(np.int64)

    std = step(rb)
    np.testing.assert_array_equal(std, fused(ctx, dev, pairs, n_batch, seed, pair_id_offset=5).astype(np.int64))
    rolled = np.roll(rb.reshape(4, 2), 1, axis=0).ravel()             # same rectangle, corners listed from another start
    np.testing.assert_array_equal(step(rolled), std)
    quad = rb.copy() + np.array([0.4, -0.2, 0.9, 0.1, 0.3, 0.5, -0.1, 0.2], np.float32)          # sheared and off-centre
    got = step(quad)
    for g in range(0, pairs.size, 3):
        d_z = dev.zeros(3 * n_batch, np.float32)
        ctx.fused_normals(seed, 5 + g, 0, n_batch, 3, d_z, n_batch); ctx.synchronize()
        z = np.zeros((5, n_batch), np.float32); z[:3] = dev.get(d_z).reshape(3, n_batch)
        k, _ = oracle.mc_thread(quad, poses[g], sds[g], pos[g], 0, z, n_batch, n_batch, bins, acc)
        assert k == got[g], g
    assert (got != std).any()


def test_stream_id_overflow_is_rejected(ctx, dev, satmc, workloads):
    pairs = workloads.dataset_pairs(8, seed=3)
    rb, poses, sds, pi, si, pos = workloads.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.zeros(3, np.float32)
    d = [dev.put(a) for a in (rb, poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
    d_cps = dev.zeros(8, np.float32); d_done = dev.zeros(8, np.int32)
    with pytest.raises(satmc.SatmcError, match="32 bits"):
        ctx.mc_step(d[0], d[1], 8, d[2], 8, d[3], d[4], d[5], d_cps, d[6], d[7], 4, d_done, 0, 100, 100, 8, 1, 0xfffffffc)


def test_item_sample_cap(ctx, dev, workloads):
    """An item's hits are summed in 32 bits (per-lane counters, the warp total, the sweep's shared counters): the planner
    must never cut an item longer than 2^31 samples, for any kind of call; and a pair that always collides counts every
    one of 2^33 + 5 samples."""
    for kind in range(4):
        for n_pairs, n_samples in ((1, 1 << 40), (100_000, 1 << 40), (100_000, (1 << 32) + 3), (3, (1 << 33) - 1)):
            chunk, n_chunks = ctx.plan_debug(kind, n_pairs, n_samples)
            assert chunk <= 1 << 31 and chunk * n_chunks >= n_samples, (kind, n_pairs, n_samples, chunk, n_chunks)
    one = workloads.cfg2_pair()
    one["rx"] = 0.0; one["ry"] = 0.0; one["sd_x"] = 0.01; one["sd_y"] = 0.01; one["sd_theta"] = 0.01
    n = (1 << 33) + 5
    assert int(fused(ctx, dev, one, n, 3)[0]) == n


def test_planner_balances_small_and_long_single_pair_calls(ctx):
    """Planner properties the measured latencies rest on: a cfg 2 call (one pair x 1e6) is cut so
    that no SM gets two blocks while others get one; a clamped chunking fills whole rounds of the resident warps; a long
    single-pair call (a cfg 4 share) gets many items per warp, none above the fused cap of 2^18 samples."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    resident_warps = sms * 2 * 8
    chunk, n_chunks = ctx.plan_debug(0, 1, 1_000_000)
    blocks = -(-n_chunks // 8)
    assert chunk % 128 == 0 and chunk * n_chunks >= 1_000_000
    assert blocks <= sms or blocks % sms == 0, (chunk, n_chunks, blocks, sms)
    chunk, n_chunks = ctx.plan_debug(0, 1, 8_000_000)
    assert n_chunks <= resident_warps, (chunk, n_chunks)             # one round, not one and a half
    chunk, n_chunks = ctx.plan_debug(0, 1, 12_500_000_000)
    assert chunk <= 1 << 18 and n_chunks >= 16 * resident_warps, (chunk, n_chunks)


def test_c_client_of_the_group_api(satmc, tmp_path):
    """tests/c/group_example.c (plain C99): satmc_group_create over every GPU of the box, both shard modes through host
    buffers, the counts equal satmc_count_fused_host on one device."""
    libdir = os.path.dirname(satmc.LIB_PATH)
    exe = str(tmp_path / "group_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", satmc.INCLUDE_DIR,
                           os.path.join(ROOT, "tests", "c", "group_example.c"), "-o", exe, "-L", libdir, "-lsatmc", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical" in r.stdout
