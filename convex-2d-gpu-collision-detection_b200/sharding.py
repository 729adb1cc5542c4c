"""Multi-GPU sharding of the Monte Carlo SAT path: one process per GPU, torch.distributed for plumbing.

The path has no data dependence between units, so it shards two ways:

The split itself lives behind the C ABI (``satmc_group_*`` in include/satmc.h: NCCL all-reduce / all-gather inside
the library); this module restates its slice arithmetic for the CPU tests (gloo, world size 2) and offers the same
two decompositions over any collective callable.

* by pair (BASELINE cfg 3 / 5): rank r owns a contiguous slice of the pair array; Philox stream ids stay
  global through ``pair_id_offset``; no data-path collective (results are gathered only if the caller asks).
* by sample range (cfg 4): rank r owns sample indices [lo_r, hi_r) of every pair (``sample_offset``); one
  all-reduce(SUM) of the 64-bit hit counters is the path's only exchange step.

Either way the totals are bit-identical to a single-GPU run: the normals of sample s of pair p depend only on
(seed, p, s).  Sample ranges are cut at multiples of 4 (the sampler's group size) so no group is split.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def pair_slice(n_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Slice [lo, hi) of the pair range owned by `rank`: equal chunks of ceil(n / world) -- the layout of the in-place
    all-gather inside ``satmc_group_count_fused`` (``satmc_shard_range(SATMC_SHARD_BY_PAIR)``, same arithmetic)."""
    chunk = -(-n_pairs // world)
    return min(rank * chunk, n_pairs), min((rank + 1) * chunk, n_pairs)


def sample_slice(n_samples: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """Slice [lo, hi) of the sample range owned by `rank`, cut at multiples of `align` (the sampler's group size);
    ``satmc_shard_range(SATMC_SHARD_BY_SAMPLE_RANGE)``."""
    blocks = (n_samples + align - 1) // align
    base, rem = divmod(blocks, world)
    lo_b = rank * base + min(rank, rem)
    hi_b = lo_b + base + (1 if rank < rem else 0)
    return min(lo_b * align, n_samples), min(hi_b * align, n_samples)


def interleaved_indices(n_pairs: int, rank: int, world: int) -> np.ndarray:
    """Round-robin assignment for adaptive (z-test) workloads, where work per pair varies ~400x
    (``satmc_shard_range(SATMC_SHARD_INTERLEAVED)``; what ``satmc_group_adaptive_run_host`` does)."""
    return np.arange(rank, n_pairs, world)


def count_by_pair(counter: Callable, pairs: np.ndarray, n_samples: int, seed: int, rank: int, world: int,
                  gather: Optional[Callable] = None):
    """counter(pairs_slice, n_samples, seed, sample_offset, pair_id_offset) -> uint64 hits of the slice.
    Returns (lo, hi, hits_slice) or, with `gather` (an all-gather over ranks of a numpy array), the full vector."""
    lo, hi = pair_slice(pairs.size, rank, world)
    hits = counter(pairs[lo:hi], n_samples, seed, 0, lo)
    if gather is None:
        return lo, hi, hits
    return np.concatenate(gather(hits))


def count_by_sample_range(counter: Callable, pairs: np.ndarray, n_samples: int, seed: int, rank: int, world: int,
                          all_reduce_sum: Callable):
    """Every rank counts its sample range of every pair; `all_reduce_sum` sums the uint64 counters over ranks."""
    lo, hi = sample_slice(n_samples, rank, world)
    hits = counter(pairs, hi - lo, seed, lo, 0)
    return all_reduce_sum(hits)


def torch_all_reduce_sum(hits: np.ndarray) -> np.ndarray:
    """all-reduce(SUM) of uint64 counters through torch.distributed (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(hits.view(np.int64).copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy().view(np.uint64)


def torch_all_gather(hits: np.ndarray):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, hits)
    return out
