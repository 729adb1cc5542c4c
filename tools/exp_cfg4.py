import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
satmc = importlib.import_module("convex-2d-gpu-collision-detection_b200")
wl = importlib.import_module("convex-2d-gpu-collision-detection_b200.workloads")
ctx = satmc.Context(0, torch.cuda.current_stream().cuda_stream)
put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.float32)).cuda()
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
one = wl.cfg2_pair()
rep = np.repeat(one, 100_000)
d_rep = put(rep); d_h = torch.zeros(100_000, dtype=torch.int64, device="cuda")
t = timeit(lambda: ctx.count_fused(d_rep, 100_000, 10_000, 7, d_h)); print("cfg2 pair x1e5, 1e4 samples:", 1e9 / t / 1e6, "Gtests/s")
ds = wl.dataset_pairs(100_000, 3); d_ds = put(ds)
t = timeit(lambda: ctx.count_fused(d_ds, 100_000, 10_000, 7, d_h)); print("dataset pairs:", 1e9 / t / 1e6)
d_one = put(one); d_h1 = torch.zeros(1, dtype=torch.int64, device="cuda")
for n in (10**9, 10**10):
    t = timeit(lambda: ctx.count_fused(d_one, 1, n, 7, d_h1), reps=3); print("1 pair x", n, ":", n / t / 1e6)
# 8 copies of the pair x 1.25e9: chunked, block-uniform
d_8 = put(np.repeat(one, 8)); d_h8 = torch.zeros(8, dtype=torch.int64, device="cuda")
t = timeit(lambda: ctx.count_fused(d_8, 8, 1_250_000_000, 7, d_h8), reps=3); print("8 pairs x 1.25e9:", 1e10 / t / 1e6)
# far-apart pair (no undecided samples) single pair 1e10
far = one.copy(); far["rx"] = 50.0
d_far = put(far)
t = timeit(lambda: ctx.count_fused(d_far, 1, 10**10, 7, d_h1), reps=3); print("far pair x 1e10:", 1e10 / t / 1e6)
print("exact evals", ctx.exact_evals())
