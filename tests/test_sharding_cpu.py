"""N>1 host logic on CPU: shard arithmetic + the collectives, world_size 2 over gloo.

The compute inside each rank is the oracle's CPU restatement of the fused path (tests may use the oracle);
what is under test is that the sharding parameters (pair_id_offset, sample_offset, cut alignment) reproduce
the single-process totals exactly, which is the property the GPU path relies on (tests/test_gpu_parity.py
checks the same invariance on the device)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "convex-2d-gpu-collision-detection_b200"


def test_slices_partition_the_ranges():
    """The Python restatement and the library's own satmc_shard_range (host arithmetic, callable without a GPU) agree and
    partition the ranges."""
    sh = importlib.import_module(PKG + ".sharding")
    mod = importlib.import_module(PKG)
    for n in (0, 1, 7, 8, 100_000, 100_003):
        for world in (1, 2, 3, 8):
            cuts = [sh.pair_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            chunk = -(-n // world)
            assert all(l == min(r * chunk, n) for r, (l, _) in enumerate(cuts))       # the in-place all-gather layout
            assert cuts == [mod.shard_range(mod.SHARD_BY_PAIR, n, world, r) for r in range(world)]
            assert [sh.sample_slice(n, r, world) for r in range(world)] == \
                [mod.shard_range(mod.SHARD_BY_SAMPLE_RANGE, n, world, r) for r in range(world)]
            for r in range(world):
                first, count = mod.shard_range(mod.SHARD_INTERLEAVED, n, world, r)
                idx = sh.interleaved_indices(n, r, world)
                assert count == idx.size and (count == 0 or first == idx[0])
            sc = [sh.sample_slice(n, r, world) for r in range(world)]
            assert sc[0][0] == 0 and sc[-1][1] == n and all(a[1] == b[0] for a, b in zip(sc, sc[1:]))
            assert all(l % 4 == 0 or l == n for l, _ in sc)
            idx = np.concatenate([sh.interleaved_indices(n, r, world) for r in range(world)])
            assert sorted(idx.tolist()) == list(range(n))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.binding import Oracle
    sh = importlib.import_module(PKG + ".sharding")
    wl = importlib.import_module(PKG + ".workloads")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    orc = Oracle()

    def counter(pairs, n, seed, sample_offset, pair_id_offset):
        return orc.count_fused_batch(pairs, n, seed, sample_offset=sample_offset, pair_id_offset=pair_id_offset, threads=1)

    pairs = wl.dataset_pairs(41, seed=13, shape_variance=True)
    pairs["sd_w"][::2] = 0; pairs["sd_h"][::2] = 0
    n, seed = 3001, 77
    by_pair = sh.count_by_pair(counter, pairs, n, seed, rank, world, gather=sh.torch_all_gather)
    by_range = sh.count_by_sample_range(counter, pairs, n, seed, rank, world, sh.torch_all_reduce_sum)
    # interleaved rows (the adaptive path's assignment): local row e of rank r is row r + e * world and draws that stream
    mine = sh.interleaved_indices(pairs.size, rank, world)
    part = np.array([counter(pairs[i:i + 1], n, seed, 0, int(i))[0] for i in mine], dtype=np.uint64)
    parts = sh.torch_all_gather(part)
    inter = np.zeros(pairs.size, np.uint64)
    for r, p in enumerate(parts):
        inter[r::world] = p
    dist.barrier()
    if rank == 0:
        q.put((by_pair, by_range, inter))
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process_counts():
    from oracle.binding import Oracle
    wl = importlib.import_module(PKG + ".workloads")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    by_pair, by_range, inter = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pairs = wl.dataset_pairs(41, seed=13, shape_variance=True)
    pairs["sd_w"][::2] = 0; pairs["sd_h"][::2] = 0
    want = Oracle().count_fused_batch(pairs, 3001, 77, threads=2)
    np.testing.assert_array_equal(by_pair, want)
    np.testing.assert_array_equal(by_range, want)
    np.testing.assert_array_equal(inter, want)
    assert want.sum() > 0


def test_group_entry_points_fail_loudly_without_a_gpu():
    """The group API has no CPU path either: creating a group without a device is an error, not a fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mod = importlib.import_module(PKG)
    with pytest.raises(mod.SatmcError):
        mod.Group(devices=[0])
    with pytest.raises(mod.SatmcError):
        mod.Group.from_rank(None, 1, 0, 0)
    with pytest.raises(mod.SatmcError):
        mod.shard_range(7, 10, 2, 0)
