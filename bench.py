#!/usr/bin/env python
"""bench.py -- headline benchmark of the Monte Carlo SAT collision-probability path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1: under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores

Metric (BASELINE.json): SAT pair-tests/s (and pair-probabilities/s).  A "step" is one pass of the hot
path over one batch of synthetic input.  Default workload = BASELINE config 3, the generate_dataset
batch: 1e5 rectangle pairs x 1e4 Monte Carlo samples per GPU (fused Philox sampler + SAT), weak
scaling: every rank owns its own batch, no data-path collective.  `--workload cfg4` is the single pair
sharded by sample range with one NCCL all-reduce of the 64-bit hit count; `--workload cfg5` the
variance sweep (64 000 rows x 1e5 samples).

One JSON line is printed by rank 0 (see the task contract for the keys).  Timing: CUDA events on the
launching stream around every step, L2 flushed (256 MiB write) before each timed step, barrier +
synchronize around the timed region, max over ranks.  `e2e` goes through the C ABI's host-buffer entry
point (pinned H2D of the pair descriptors + kernel + D2H of the hit counts inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "convex-2d-gpu-collision-detection_b200"

METRIC = "SAT pair-tests/sec"
UNIT = "tests/s"
# Issue cost of one test in the fused 3-DoF loop, read off the SASS of k_count: per 4-sample
# group 337 instructions, 53 of them IMAD.WIDE (the loop-invariant Philox products are hoisted).  On sm_100a an IMAD.WIDE
# holds the warp scheduler's issue port for ~4 cycles and does not overlap with FP32 issue (tools/ubench.cu:
# "IMAD.WIDE+LOP3+4 FFMA" = 9.4 clk, purely additive; profiles/r1_ubench.log), so a group costs 284 + 53*4 = 496 issue
# slots = 124 lane-slots per test.
ISSUE_SLOTS_PER_TEST = 124.0
INSTR_PER_TEST = 337.0 / 4.0     # plain count, every instruction one slot: what ncu's issue-active measures
# FMA-pipe view of the same loop: 124 FP32 + 53 IMAD.WIDE x 4 = 336 FFMA-equivalent slots per group -> 84 per test
FMA_SLOTS_PER_TEST = 84.0
SURVEY_I_FMA_W8 = 247.0      # 8-axis kernel, 3-DoF
SURVEY_I_FMA_W4 = 163.0      # 4-axis kernel (+ exact fallback), 3-DoF


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def workload_pairs(wl, name, rank):
    if name == "cfg3":
        return wl.dataset_pairs(100_000, seed=3 + 1000 * rank), 10_000
    if name == "cfg5":
        return wl.variance_sweep_pairs(1000, seed=5 + 1000 * rank), 100_000
    if name == "cfg4":
        return wl.cfg2_pair(), 20_000_000_000       # per-rank share of the sample range (weak scaling)
    if name == "cfg2":
        return wl.cfg2_pair(), 1_000_000
    raise SystemExit(f"unknown workload {name}")


WORKLOAD_DESC = {
    "cfg3": "cfg3 generate_dataset batch: 1e5 pairs x 1e4 samples per GPU, fused sampler, sharded by pair",
    "cfg5": "cfg5 variance sweep slice: 64000 (pair,covariance) rows x 1e5 samples per GPU, fused sampler",
    "cfg4": "cfg4 single pair, 2e10 samples per GPU, sharded by sample range + NCCL all-reduce of the count",
    "cfg2": "cfg2 single pair x 1e6 samples",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(mod, nbytes, dtype):
    """numpy view of pinned host memory from satmc_host_alloc."""
    lib = mod.load_library()
    p = ctypes.c_void_p()
    if lib.satmc_host_alloc(ctypes.byref(p), nbytes) != 0:
        raise RuntimeError("satmc_host_alloc failed")
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype), p


def cpu_baseline(pairs, n_samples, seed, target_s=10.0, threads=0):
    """The oracle port (sampler + reference geometry restated in C) on the host cores, bounded sample."""
    from oracle.binding import Oracle
    orc = Oracle()
    cores = orc.hardware_threads() if threads <= 0 else threads
    n_p = min(pairs.size, 8 * cores)
    ns = min(n_samples, 2000)
    t0 = time.perf_counter()
    orc.count_fused_batch(pairs[:n_p], ns, seed, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = n_p * ns / dt                                  # calibration
    want = rate * target_s
    ns = min(n_samples, 10_000)
    n_p = int(min(pairs.size, max(cores, want // ns)))
    n_p = max(cores, (n_p // cores) * cores) if pairs.size >= cores else pairs.size
    n_p = min(n_p, pairs.size)
    if pairs.size == 1:                                   # single-pair workloads: bound the sample count instead
        n_p, ns = 1, int(min(n_samples, max(10_000, want)))
        cores = 1
    t0 = time.perf_counter()
    orc.count_fused_batch(pairs[:n_p], ns, seed, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n_p * ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_p} pairs x {ns} samples of the same workload, oracle/sat_oracle.c (Philox+Box-Muller sampler, "
                      f"8-axis reference SAT), {cores} threads, {dt:.1f} s"}, n_p, ns


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (the reference has no CPU implementation
    of its own -- convex_collide is __device__-only and SAT.py is absent upstream -- so this is the oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mod = importlib.import_module(PKG)
    wl = importlib.import_module(PKG + ".workloads")
    pairs, n_samples = workload_pairs(wl, args.workload, 0)
    base, n_p, ns = cpu_baseline(pairs, n_samples, 7, target_s=2.0)
    from oracle.binding import Oracle
    orc = Oracle()
    cores = base["cores"]
    for _ in range(args.warmup):
        orc.count_fused_batch(pairs[:n_p], ns, 7, threads=cores)
    t0 = time.perf_counter()
    for s in range(args.steps):
        orc.count_fused_batch(pairs[:n_p], ns, 7 + s, threads=cores)
    dt = time.perf_counter() - t0
    value = n_p * ns * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "step_sample": f"{n_p} pairs x {ns} samples per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_p} pairs x {ns} samples per step, {args.steps} steps, oracle/sat_oracle.c"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="satmc", choices=["satmc", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (streamed roofline, reference GPU kernel)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this library has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mod = importlib.import_module(PKG)
    wl = importlib.import_module(PKG + ".workloads")
    stream = torch.cuda.current_stream().cuda_stream
    ctx = mod.Context(local, stream)

    pairs, n_samples = workload_pairs(wl, args.workload, rank)
    n_pairs = pairs.size
    pair_id_offset = rank * n_pairs if args.workload != "cfg4" else 0
    sample_offset = rank * n_samples if args.workload == "cfg4" else 0
    d_pairs = torch.from_numpy(np.ascontiguousarray(pairs).view(np.float32)).cuda()
    d_hits = torch.zeros(n_pairs, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")           # > 126 MB L2
    seed = 20261018

    def step(s):
        ctx.count_fused(d_pairs, n_pairs, n_samples, seed + s, d_hits, sample_offset=sample_offset,
                        pair_id_offset=pair_id_offset)
        if args.workload == "cfg4" and world > 1:
            dist.all_reduce(d_hits)                                           # the path's one exchange step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    t_sampler = time.perf_counter()
    if rank == 0:
        sampler.start()                        # samples every 50 ms through warm-up, the timed region and the e2e loop
    for s in range(args.warmup):
        flush.fill_(s & 0xff)
        step(s)
    barrier()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xff)                                                  # L2 flush, outside the event pair
        ev[s][0].record()
        step(args.warmup + s)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    hits_total = int(d_hits.sum().item())

    # e2e: host buffers through the C ABI (pinned H2D + kernel + D2H inside the timed region)
    h_pairs, p1 = pinned_array(mod, n_pairs * 48, np.uint8)
    h_hits, p2 = pinned_array(mod, n_pairs * 8, np.uint64)
    h_pairs[:] = np.frombuffer(pairs.tobytes(), dtype=np.uint8)
    h_pairs_struct = h_pairs.view(mod.PAIR_DTYPE)
    for s in range(2):
        ctx.count_fused_host(h_pairs_struct, n_samples, seed + s, sample_offset, pair_id_offset, out=h_hits)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        ctx.count_fused_host(h_pairs_struct, n_samples, seed + 100 + s, sample_offset, pair_id_offset, out=h_hits)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_kernel_ms = ctx.last_kernel_ms()
    while rank == 0 and (world == 1 or args.workload != "cfg4") and time.perf_counter() - t_sampler < 0.8:   # short runs: keep the load on until
        step(0)                                                                # nvidia-smi has had time to sample it
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s * 1e3, t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, wall_ms_max = (float(x) for x in t.cpu())

    if rank == 0:
        tests_per_step = float(n_pairs) * float(n_samples) * world
        value = tests_per_step * args.steps / (dev_ms_max * 1e-3)
        per_gpu = value / world
        hbm_peak, sm_mhz, src = measured_peaks()
        fma_peak = 148 * 128 * sm_mhz * 1e6                                    # FP32 lane-slots/s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "pairs_per_gpu": n_pairs, "samples_per_pair": n_samples,
                       "path": "fused Philox4x32-10 + Box-Muller sampler -> screened SAT (exact 8-axis fallback)",
                       "sharding": "by sample range + NCCL all-reduce" if args.workload == "cfg4" else "by pair, no collective",
                       "l2": "flushed (256 MiB write) before every timed step", "timing": "CUDA events per step, max over ranks"},
            "pair_probabilities_per_s": n_pairs * world * args.steps / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": tests_per_step * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_pairs * 48, "d2h_bytes_per_step": n_pairs * 8,
                    "api": "satmc_count_fused_host (C ABI, pinned host buffers)", "kernel_ms_last_step": e2e_kernel_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32-issue", "kernel": "satmc::k_count<DirectSrc,false> (fused, 3-DoF loop)",
                "achieved": per_gpu * ISSUE_SLOTS_PER_TEST / 1e12, "peak": fma_peak / 1e12, "unit": "T lane-issue-slots/s",
                "frac": per_gpu * ISSUE_SLOTS_PER_TEST / fma_peak, "traffic": 4.86e6,
                "peak_source": f"148 SM x 4 schedulers x 32 lanes x sm_max_mhz {sm_mhz:.0f} from {src} (= the FP32 lane peak)",
                "alg_units": "124 issue slots per test: per 4-sample group 284 single-slot instructions + 53 IMAD.WIDE x 4 slots "
                             "(IMAD.WIDE blocks issue ~4 clk on sm_100a, profiles/r1_ubench.log)",
                "frac_issue_plain": per_gpu * INSTR_PER_TEST / fma_peak,
                "frac_fma_pipe": per_gpu * FMA_SLOTS_PER_TEST / fma_peak,
                "frac_vs_survey_w8_model": per_gpu * SURVEY_I_FMA_W8 / fma_peak,
                "frac_vs_survey_w4_model": per_gpu * SURVEY_I_FMA_W4 / fma_peak,
                "note": "frac = issue-slot utilisation on this kernel's own SASS count with IMAD.WIDE weighted 4; frac_issue_plain "
                        "weights every instruction 1 (= ncu's issue-active); frac_fma_pipe counts only FMA-pipe work "
                        "(124 FP32 + 53 IMAD.WIDE x4 per group); the last two are against SURVEY.md 8(d)'s instruction models "
                        "(247 / 163 FMA-pipe instructions per test) and exceed 1 because the screening pass needs ~46",
            },
            "hits_checksum": hits_total,
        }
        if world == 1 and not args.no_extras and args.workload == "cfg3":
            line["extras"] = extras(ctx, mod, wl, torch, hbm_peak, src)
        if world == 1 and not args.no_cpu_baseline:                # the CPU baseline is an N=1 measurement
            line["cpu_baseline"], _, _ = cpu_baseline(pairs, n_samples, seed)
        print(json.dumps(line), flush=True)
    lib = mod.load_library()
    lib.satmc_host_free(p1); lib.satmc_host_free(p2)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(ctx, mod, wl, torch, hbm_peak, src):
    """Secondary measurements on rank 0: streamed-sample path against the HBM roofline, other configs, and
    the reference's own GPU kernel recompiled for sm_100a (oracle/_ref) on the same box."""
    out = {}

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    put = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.float32)).cuda()
    npairs, n = 16384, 32768
    pp = wl.dataset_pairs(npairs, 9); d_pp = put(pp); d_h = torch.zeros(npairs, dtype=torch.int64, device="cuda")
    for ndof in (5, 3):
        z = torch.randn(ndof * npairs * n, device="cuda")          # 10.7 / 6.4 GB, consumed once per pass (> L2)
        ms = timed(lambda: ctx.count_streamed(d_pp, npairs, z, npairs * n, ndof, n, d_h, z_pair_stride=n))
        gbs = ndof * 4.0 * npairs * n / ms / 1e6
        out[f"streamed_hbm_ndof{ndof}"] = {"tests_per_s": npairs * n / ms * 1e3, "ms": ms,
                                           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                        "frac": gbs / hbm_peak, "alg_bytes_per_test": 4 * ndof, "peak_source": src}}
        del z
    sweep = wl.variance_sweep_pairs(1000, 5); d_sw = put(sweep); d_hs = torch.zeros(sweep.size, dtype=torch.int64, device="cuda")
    ms = timed(lambda: ctx.count_fused(d_sw, sweep.size, 100_000, 7, d_hs), reps=3)
    out["cfg5_fused"] = {"tests_per_s": sweep.size * 1e5 / ms * 1e3, "ms": ms, "rows": int(sweep.size), "samples": 100_000}
    grid = np.array([0.01, 0.05, 0.15, 0.3]); vx, vy, vt = np.meshgrid(grid, grid, grid, indexing="ij")
    sig = np.sqrt(np.stack([vx.ravel(), vy.ravel(), vt.ravel()], 1)).astype(np.float32)
    base = wl.dataset_pairs(1000, 5); d_base = put(base); d_sig = torch.from_numpy(sig.ravel()).cuda()
    ms = timed(lambda: ctx.count_fused_sweep(d_base, base.size, d_sig, 64, 100_000, 7, d_hs), reps=3)
    out["cfg5_fused_sweep_common_random_numbers"] = {"tests_per_s": base.size * 64 * 1e5 / ms * 1e3, "ms": ms, "rows": int(base.size * 64),
                                                     "samples": 100_000, "what": "satmc_count_fused_sweep: 64 covariance settings per pair "
                                                     "share one Philox stream (the sampler runs once per sample, not once per setting)"}
    zb = torch.randn(3 * 100_000, device="cuda")
    ms = timed(lambda: ctx.count_streamed(d_sw, sweep.size, zb, 100_000, 3, 100_000, d_hs), reps=3)
    out["cfg5_streamed_shared_bank"] = {"tests_per_s": sweep.size * 1e5 / ms * 1e3, "ms": ms}
    one = put(wl.cfg2_pair()); d_h1 = torch.zeros(1, dtype=torch.int64, device="cuda")
    ms = timed(lambda: ctx.count_fused(one, 1, 1_000_000, 7, d_h1), reps=20, warm=5)
    out["cfg2_fused_1pair_1e6"] = {"tests_per_s": 1e6 / ms * 1e3, "ms": ms}
    # general convex polygons: the cfg 3 rectangles given as 4-vertex polygons, and octagons of similar size
    pr = wl.dataset_pairs(100_000, 3)
    d_hp = torch.zeros(pr.size, dtype=torch.int64, device="cuda")

    def rect_verts(w, h):
        return np.stack([-w / 2, -h / 2, w / 2, -h / 2, w / 2, h / 2, -w / 2, h / 2], 1).reshape(-1, 4, 2).astype(np.float32)

    def octagons(w, h):
        a = 2 * np.pi * np.arange(8) / 8
        return np.stack([0.5 * w[:, None] * np.cos(a), 0.5 * h[:, None] * np.sin(a)], 2).astype(np.float32)
    for name, rv, ov in (("polygons_4x4_cfg3_rectangles", rect_verts(pr["rw"], pr["rh"]), rect_verts(pr["ow"], pr["oh"])),
                         ("polygons_8x8_octagons", octagons(pr["rw"], pr["rh"]), octagons(pr["ow"], pr["oh"]))):
        ppoly = mod.make_poly_pairs(list(rv), list(ov), pr["rx"], pr["ry"], pr["rtheta"], pr["sd_x"], pr["sd_y"], pr["sd_theta"])
        d_poly = torch.from_numpy(np.ascontiguousarray(ppoly).view(np.uint8).view(np.float32)).cuda()
        ms = timed(lambda: ctx.count_fused_polygons(d_poly, ppoly.size, 10_000, 7, d_hp), reps=3)
        out[name] = {"tests_per_s": ppoly.size * 1e4 / ms * 1e3, "ms": ms, "mean_p": float(d_hp.sum().item()) / (ppoly.size * 1e4),
                     "what": "satmc_count_fused_polygons: circle-based screening pass, undecided samples compacted per warp, exact polygon SAT"}
    # the adaptive z-test batch (what generate_dataset / compute_collision_probability do per file)
    pairs = wl.dataset_pairs(100_000, 3)
    rb, poses, sds, pi, si, pos = wl.reference_tables(pairs)
    bins = np.array([0, 0.01, 0.1, 1.0], np.float32); acc = np.array([1e-4, 1e-3, 1e-2], np.float32)
    max_samples = 1_020_000
    dd = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rb, poses.ravel(), sds.ravel(), pi, si, pos.ravel(), bins, acc)]
    d_cp = torch.zeros(pairs.size, device="cuda")

    def adaptive():
        return ctx.adaptive_run(dd[0], dd[1], pairs.size, dd[2], pairs.size, dd[3], dd[4], dd[5], pairs.size, dd[6], dd[7], 4,
                                max_samples, 1000, 20000, 100000, 7, d_cp)
    adaptive(); torch.cuda.synchronize()
    t0 = time.perf_counter(); iters, drawn = adaptive(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out["adaptive_batch"] = {"pairs": int(pairs.size), "max_samples": max_samples, "ms": dt * 1e3, "iterations": iters,
                             "samples_drawn": drawn, "tests_per_s": drawn / dt, "pair_probabilities_per_s": pairs.size / dt,
                             "what": "satmc_adaptive_run: schedule 1000/20000/100000, bins 0|0.01|0.1|1, accuracy 1e-4|1e-3|1e-2, wall clock"}
    try:
        from oracle.binding import RefGpu
        ref = RefGpu()
        _, ms, drawn_ref = ref.adaptive_batch(rb, poses, sds, pi, si, pos, bins, acc, max_samples, 3)
        out["reference_adaptive_batch"] = {"ms": ms, "samples_drawn": drawn_ref, "tests_per_s": drawn_ref / ms * 1e3,
                                           "pair_probabilities_per_s": pairs.size / ms * 1e3,
                                           "what": "the reference's loop (its kernel + thrust::count + thrust::sort_by_key + tail copies) "
                                                   "recompiled for sm_100a, same batch and schedule, wall clock"}
        ref.mc_time(rb, poses, sds, pi, si, pos, 1000, 2, 1)
        ms, _ = ref.mc_time(rb, poses, sds, pi, si, pos, 1000, 10, 1)
        out["reference_gpu_kernel_cfg3"] = {"tests_per_s": 1e9 / ms * 1e3, "ms": ms,
                                            "what": "unmodified reference kernel (ztest.cu:106-166) recompiled for sm_100a, "
                                                    "1e5 pairs x 10 launches of n_batch=1000, CUDA events"}
    except Exception as e:                                           # the oracle build is optional on the box
        out["reference_gpu_kernel_cfg3"] = {"unavailable": str(e)[:200]}
    return out


if __name__ == "__main__":
    main()
