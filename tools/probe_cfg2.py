#!/usr/bin/env python
"""Development probe (GPU box): device time of small calls (BASELINE cfg 2 and neighbours) against the sample count,
and an empty-kernel floor, to separate launch overhead from work."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")


def med_us(fn, s, reps=200):
    for _ in range(20):
        fn()
    s.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); fn(); b.record(s); s.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def main():
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ctx = satmc.Context(0, s.cuda_stream)
        one = torch.from_numpy(np.ascontiguousarray(wl.cfg2_pair()).view(np.float32)).cuda()
        d_h = torch.zeros(1, dtype=torch.int64, device="cuda")
        x = torch.zeros(32, device="cuda")
        print("tiny torch kernel (fill_ of 32 floats): med %.2f us min %.2f us" % med_us(lambda: x.fill_(1.0), s))
        for n in (128, 4096, 100_000, 200_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 6_000_000, 8_000_000, 12_000_000, 16_000_000, 30_000_000, 60_000_000):
            chunk, n_chunks = ctx.plan_debug(0, 1, n)
            m, lo = med_us(lambda: ctx.count_fused(one, 1, n, 7, d_h), s)
            print(f"1 pair x {n:>9d}: med {m:7.2f} us  min {lo:7.2f} us   chunk {chunk} x {n_chunks} items   {n / m / 1e3:8.2f} Gtests/s")
        ctx.close()


if __name__ == "__main__":
    main()
