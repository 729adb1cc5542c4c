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
    sh = importlib.import_module(PKG + ".sharding")
    for n in (0, 1, 7, 8, 100_000, 100_003):
        for world in (1, 2, 3, 8):
            cuts = [sh.pair_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
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
    dist.barrier()
    if rank == 0:
        q.put((by_pair, by_range))
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
    by_pair, by_range = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pairs = wl.dataset_pairs(41, seed=13, shape_variance=True)
    pairs["sd_w"][::2] = 0; pairs["sd_h"][::2] = 0
    want = Oracle().count_fused_batch(pairs, 3001, 77, threads=2)
    np.testing.assert_array_equal(by_pair, want)
    np.testing.assert_array_equal(by_range, want)
    assert want.sum() > 0
