#!/usr/bin/env python
"""Development probe for ncu: a handful of cfg 2 calls (1 pair x 1e6 samples) so that gpu__time_duration of the kernel
alone can be read off the launch list."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
one = torch.from_numpy(np.ascontiguousarray(wl.cfg2_pair()).view(np.float32)).cuda()
d_h = torch.zeros(1, dtype=torch.int64, device="cuda")
for n in (128, 100_000, 1_000_000, 1_000_000, 1_000_000, 4_000_000):
    ctx.count_fused(one, 1, n, 7, d_h)
torch.cuda.synchronize()
print(int(d_h.item()))
