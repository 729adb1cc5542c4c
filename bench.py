#!/usr/bin/env python
"""bench.py -- headline benchmark of the Monte Carlo SAT collision-probability path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N>1: under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores

Metric (BASELINE.json): SAT pair-tests/s (and pair-probabilities/s).  A "step" is one pass of the hot
path over one batch of synthetic input.  Headline workload = BASELINE config 3, the generate_dataset
batch: 1e5 rectangle pairs x 1e4 Monte Carlo samples per GPU (fused Philox sampler + SAT), weak
scaling: every rank owns its own batch, no data-path collective.

Beside the headline the same run times, at every N, the STRONG-scaled forms of the two sharded configs
through the library's own group entry points (satmc_group_*: the collective is NCCL inside libsatmc.so,
not torch.distributed):
    strong.cfg3   1e5 pairs in TOTAL, sharded by pair, counters all-gathered to every rank
    strong.cfg4   one pair, 1e11 samples in TOTAL, sharded by sample range, ONE ncclAllReduce(u64) of the
                  count, its time reported separately (allreduce_us)
    sharding_check  the N-rank count of a 1e9-sample range equals one GPU's count of the same range, bit for bit
and, on rank 0 at N >= 2, runs the drop-in `ztest` program with --gpus N and --gpus 1 on the same rows
(its output must not depend on the GPU count).

One JSON line is printed by rank 0.  Timing: CUDA events on the launching stream around every step, L2
flushed (256 MiB write) before each timed step, barrier + synchronize around the timed region, max over
ranks.  `e2e` goes through the C ABI's host-buffer entry point (pinned H2D of the pair descriptors +
kernel + D2H of the hit counts inside the timed region).  The per-test instruction constants of the
roofline object are read from profiles/r2_sass_k_count_hotloop.json (tools/sass_count.py).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

# stdout carries exactly one JSON line.  NCCL prints its "NCCL version ..." banner to fd 1 from C when the first
# communicator comes up, and child programs inherit fd 1: the real stdout is set aside for emit() and fd 1 is pointed
# at stderr for everything else.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    fd = 1 if _JSON_FD is None else _JSON_FD
    while data:
        data = data[os.write(fd, data):]

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "convex-2d-gpu-collision-detection_b200"

METRIC = "SAT pair-tests/sec"
UNIT = "tests/s"
SURVEY_I_FMA_W8 = 247.0      # 8-axis kernel, 3-DoF
SURVEY_I_FMA_W4 = 163.0      # 4-axis kernel (+ exact fallback), 3-DoF
REFERENCE_THREADS = 16       # --impl reference and cpu_baseline use this many host threads (or all, if the box has fewer)


def sass_constants():
    """Instruction mix of the fused 3-DoF hot loop (one trip = one 4-sample group per lane), from the committed artefact
    that tools/sass_count.py writes; the loop with 50..60 IMAD.WIDE is the 3-DoF one (3 Philox calls)."""
    path = os.path.join(ROOT, "profiles", "r2_sass_k_count_hotloop.json")
    out = {"instr_per_group": 302.0, "imad_wide_per_group": 53.0, "mufu_per_group": 32.0, "fp32_per_group": 67.0, "fp32x2_per_group": 34.0,
           "source": "built-in defaults (profiles/r2_sass_k_count_hotloop.json missing)"}
    try:
        with open(path) as f:
            d = json.load(f)
        for name, k in d["kernels"].items():
            if "DirectSrc, false, false, false" not in name:
                continue
            for L in k["loops"]:
                if 45 <= L["imad_wide"] <= 60:
                    out = {"instr_per_group": float(L["instr"]), "imad_wide_per_group": float(L["imad_wide"]),
                           "mufu_per_group": float(L["mufu"]), "fp32_per_group": float(L["fp32"]),
                           "fp32x2_per_group": float(L.get("fp32x2", 0)),
                           "source": "profiles/r2_sass_k_count_hotloop.json (tools/sass_count.py on the shipped libsatmc.so)"}
    except (OSError, KeyError, ValueError):
        pass
    return out


def ncu_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_fused.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


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
    "cfg4": "cfg4 single pair, 2e10 samples per GPU, sharded by sample range + ncclAllReduce of the count (satmc_group_count_fused)",
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
        busy = [x for x in sm if x >= 0.6 * max(mx)] or sm          # samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(mod, nbytes, dtype):
    """numpy view of pinned host memory from satmc_host_alloc."""
    lib = mod.load_library()
    p = ctypes.c_void_p()
    if lib.satmc_host_alloc(ctypes.byref(p), nbytes) != 0:
        raise RuntimeError("satmc_host_alloc failed")
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype), p


def reference_threads(orc):
    """One thread count for every CPU number of a run: REFERENCE_THREADS, or fewer if the box has fewer."""
    return max(1, min(REFERENCE_THREADS, orc.hardware_threads(), orc.affinity_count()))


def cpu_baseline(pairs, n_samples, seed, target_s=10.0, threads=0):
    """The oracle port (sampler + reference geometry restated in C) on the host cores, bounded sample."""
    from oracle.binding import Oracle
    orc = Oracle()
    cores = reference_threads(orc) if threads <= 0 else threads
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
            "hardware_concurrency": orc.hardware_threads(), "affinity_cpus": orc.affinity_count(),
            "sample": f"{n_p} pairs x {ns} samples of the same workload, oracle/sat_oracle.c (Philox+Box-Muller sampler, "
                      f"8-axis reference SAT), {cores} threads, {dt:.1f} s"}, n_p, ns


def cpu_baselines_c1_c2(wl):
    """C1 = cfg 1 (10 000 random rectangle pairs, SAT only, repeated), C2 = cfg 2 (one pair x 1e6
    shared normals: scale -> transform -> 8-axis SAT -> count), each on 1 thread and on all the threads the arm uses."""
    from oracle.binding import Oracle
    orc = Oracle()
    cores = reference_threads(orc)
    out = {"hardware_concurrency": orc.hardware_threads(), "affinity_cpus": orc.affinity_count(), "threads_all": cores,
           "what": "oracle/sat_oracle.c, the C restatement of utils.cu:119-184 (the reference has no CPU SAT of its own)"}
    r1, r2 = wl.cfg1_rect_pairs(10_000, seed=1)
    z = wl.normal_bank(1_000_000, 5, seed=2)
    pair = wl.cfg2_pair()
    dec1 = None
    for label, th in (("1_thread", 1), ("all_threads", cores)):
        reps = 100 if th == 1 else 100 * min(cores, 8)
        orc.sat_batch_mt(r1, r2, 2, th)
        t0 = time.perf_counter(); dec = orc.sat_batch_mt(r1, r2, reps, th); dt = time.perf_counter() - t0
        dec1 = dec if dec1 is None else dec1
        out[f"C1_cfg1_sat_only_{label}"] = {"tests_per_s": 10_000 * reps / dt, "threads": th, "pairs": 10_000, "repeats": reps,
                                            "seconds": dt, "collisions": int(dec.sum()), "same_decisions": bool((dec == dec1).all())}
        reps2 = 3 if th == 1 else 3 * min(cores, 8)
        orc.count_streamed_mt(pair, z[:, :50_000], th)
        t0 = time.perf_counter()
        for _ in range(reps2):
            hits = orc.count_streamed_mt(pair, z, th)
        dt = time.perf_counter() - t0
        out[f"C2_cfg2_1pair_1e6_{label}"] = {"tests_per_s": 1e6 * reps2 / dt, "threads": th, "samples": 1_000_000, "repeats": reps2,
                                             "seconds": dt, "hits": int(hits)}
    return out


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (the reference has no CPU implementation
    of its own -- convex_collide is __device__-only and SAT.py is absent upstream -- so this is the oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
        "config": {"workload": WORKLOAD_DESC[args.workload], "step_sample": f"{n_p} pairs x {ns} samples per step",
                   "host_threads": cores, "hardware_concurrency": orc.hardware_threads(), "affinity_cpus": orc.affinity_count(),
                   "note": f"fixed at min({REFERENCE_THREADS}, available) threads so that runs on different boxes compare"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_p} pairs x {ns} samples per step, {args.steps} steps, oracle/sat_oracle.c"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="satmc", choices=["satmc", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (streamed roofline, reference GPU kernel)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (strong.cfg3 / strong.cfg4 / sharding_check)")
    ap.add_argument("--no-programs", action="store_true", help="skip the ztest --gpus N vs --gpus 1 run at N >= 2")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    claim_stdout()
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
    # the group: one rank per process; rank 0's NCCL id travels through torch.distributed (plumbing), the collectives of
    # the data path are issued by libsatmc.so itself on the stream given here
    uid = None
    if world > 1:
        t_id = torch.zeros(mod.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t_id.copy_(torch.frombuffer(bytearray(mod.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(t_id, 0)
        uid = bytes(t_id.cpu().numpy().tobytes())
    group = mod.Group.from_rank(uid, world, rank, local, stream)

    pairs, n_samples = workload_pairs(wl, args.workload, rank)
    n_pairs = pairs.size
    is_cfg4 = args.workload == "cfg4"
    pair_id_offset = rank * n_pairs if not is_cfg4 else 0
    d_pairs = torch.from_numpy(np.ascontiguousarray(pairs).view(np.float32)).cuda()
    d_hits = torch.zeros(max(n_pairs, group.hits_capacity(n_pairs)), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")           # > 126 MB L2
    seed = 20261018

    def step(s):
        if is_cfg4:      # every rank counts its range of the world * n_samples samples; one ncclAllReduce inside the call
            group.count_fused(d_pairs, n_pairs, n_samples * world, seed + s, mod.SHARD_BY_SAMPLE_RANGE, d_hits)
        else:
            ctx.count_fused(d_pairs, n_pairs, n_samples, seed + s, d_hits, pair_id_offset=pair_id_offset)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches_now():
        return ctx.launch_count + group.launch_count()

    sampler = ClockSampler(local)
    t_sampler = time.perf_counter()
    if rank == 0:
        sampler.start()                        # samples every 50 ms through warm-up, the timed region and the e2e loop
    for s in range(args.warmup):
        flush.fill_(s & 0xff)
        step(s)
    barrier()
    launches0 = launches_now()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xff)                                                  # L2 flush, outside the event pair
        ev[s][0].record()
        step(args.warmup + s)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = launches_now() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    hits_total = int(d_hits[:n_pairs].sum().item())

    # e2e: host buffers through the C ABI (pinned H2D + kernel + D2H inside the timed region)
    h_pairs, p1 = pinned_array(mod, n_pairs * 48, np.uint8)
    h_hits, p2 = pinned_array(mod, n_pairs * 8, np.uint64)
    h_pairs[:] = np.frombuffer(pairs.tobytes(), dtype=np.uint8)
    h_pairs_struct = h_pairs.view(mod.PAIR_DTYPE)

    def e2e_step(s):
        if is_cfg4:
            group.count_fused_host(h_pairs_struct, n_samples * world, seed + s, mod.SHARD_BY_SAMPLE_RANGE, out=h_hits)
        else:
            ctx.count_fused_host(h_pairs_struct, n_samples, seed + s, 0, pair_id_offset, out=h_hits)
    for s in range(2):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        e2e_step(100 + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_kernel_ms = ctx.last_kernel_ms() if not is_cfg4 else None
    while rank == 0 and not is_cfg4 and time.perf_counter() - t_sampler < 0.8:   # short runs: keep the load on until
        step(0)                                                                # nvidia-smi has had time to sample it
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s * 1e3, t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, wall_ms_max = (float(x) for x in t.cpu())

    strong = None
    if not args.no_strong:
        strong = strong_scaling(torch, dist, mod, wl, ctx, group, rank, world, flush, barrier, min(args.steps, 20))
    programs = None
    if world > 1 and not args.no_programs:
        programs = programs_gpu_count_invariance(dist, rank, world)

    if rank == 0:
        tests_per_step = float(n_pairs) * float(n_samples) * world
        value = tests_per_step * args.steps / (dev_ms_max * 1e-3)
        per_gpu = value / world
        hbm_peak, sm_mhz, src = measured_peaks()
        lane_peak = 148 * 128 * sm_mhz * 1e6                                   # lane issue slots/s = FP32 lane peak
        sc = sass_constants()
        instr_per_test = sc["instr_per_group"] / 4.0
        weighted_per_test = (sc["instr_per_group"] + 3.0 * sc["imad_wide_per_group"]) / 4.0
        wide_share = 4.0 * sc["imad_wide_per_group"] / (sc["instr_per_group"] + 3.0 * sc["imad_wide_per_group"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "pairs_per_gpu": n_pairs, "samples_per_pair": n_samples,
                       "path": "fused Philox4x32-10 + Box-Muller sampler -> screened SAT (exact 8-axis fallback)",
                       "sharding": "by sample range + ncclAllReduce inside satmc_group_count_fused" if is_cfg4 else "by pair, no collective",
                       "l2": "flushed (256 MiB write) before every timed step", "timing": "CUDA events per step, max over ranks"},
            "pair_probabilities_per_s": n_pairs * world * args.steps / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": tests_per_step * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_pairs * 48, "d2h_bytes_per_step": n_pairs * 8,
                    "api": ("satmc_group_count_fused_host" if is_cfg4 else "satmc_count_fused_host") + " (C ABI, pinned host buffers)",
                    "kernel_ms_last_step": e2e_kernel_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "issue", "kernel": "satmc::k_count<DirectSrc,false,false> (fused, 3-DoF loop)",
                "achieved": per_gpu * instr_per_test / 1e12, "peak": lane_peak / 1e12, "unit": "T lane-instr-slots/s",
                "frac": per_gpu * instr_per_test / lane_peak,
                "traffic": ncu_traffic(),
                "binding_resource": f"Philox IMAD.WIDE: {sc['imad_wide_per_group']:.0f} per 4-sample group; each holds the scheduler's issue port "
                                    f"~4 clk on sm_100a (tools/ubench.cu), i.e. {100 * wide_share:.0f} % of the slot-cycles of the loop",
                "peak_source": f"148 SM x 4 schedulers x 32 lanes x sm_max_mhz {sm_mhz:.0f} from {src} (numerically the FP32 lane peak)",
                "alg_units": f"{instr_per_test:.2f} warp-instructions per test = {sc['instr_per_group']:.0f} per 4-sample group, every instruction "
                             f"one issue slot: what ncu's smsp__issue_active measures ({sc['source']})",
                "frac_issue_weighted": per_gpu * weighted_per_test / lane_peak,
                "frac_issue_weighted_note": "IMAD.WIDE counted as 4 slots (it blocks issue for ~4 clk): how close the loop is to the bound "
                                            "its own instruction mix allows; ~1 means no headroom without fewer wide multiplies",
                "frac_fma_pipe": per_gpu * (sc["fp32_per_group"] + 2.0 * sc["fp32x2_per_group"] + 4.0 * sc["imad_wide_per_group"]) / 4.0 / lane_peak,
                "frac_fma_pipe_note": "FMA-pipe cycles: scalar FP32 1, packed FP32 (FFMA2/FMUL2, two operations per issue slot) 2, IMAD.WIDE 4",
                "frac_vs_survey_w8_model": per_gpu * SURVEY_I_FMA_W8 / lane_peak,
                "frac_vs_survey_w4_model": per_gpu * SURVEY_I_FMA_W4 / lane_peak,
                "note": "frac is plain issue-slot utilisation (warp-instructions issued / issue slots available). Since round 2 the "
                        "screening arithmetic is packed FP32: fewer issued instructions for the same work, so the kernel got faster "
                        "(3.34 -> 3.23 ms) while this utilisation figure FELL (0.68 -> 0.63); the last two are "
                        "against SURVEY.md 8(d)'s models (247 / 163 FMA-pipe instructions per test) and exceed 1 because a 22-FP-op "
                        "conservative screening test decides 99.98 % of the samples and only the rest run the exact 8-axis SAT",
            },
            "hits_checksum": hits_total,
        }
        if strong is not None:
            line["strong"] = strong["strong"]
            line["sharding_check"] = strong["sharding_check"]
        if programs is not None:
            line["programs_gpu_count_invariance"] = programs
        if world == 1 and not args.no_extras and args.workload == "cfg3":
            line["extras"] = extras(ctx, mod, wl, torch, hbm_peak, src)
        if world == 1 and not args.no_cpu_baseline:                # the CPU baselines are N=1 measurements
            line["cpu_baseline"], _, _ = cpu_baseline(pairs, n_samples, seed)
            line["cpu_baselines_c1_c2"] = cpu_baselines_c1_c2(wl)
        emit(line)
    lib = mod.load_library()
    lib.satmc_host_free(p1); lib.satmc_host_free(p2)
    group.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling(torch, dist, mod, wl, ctx, group, rank, world, flush, barrier, steps):
    """The regime where launch latency, host<->device copies and the collective can lose: TOTAL work fixed as N grows.
    Everything goes through satmc_group_* (NCCL inside libsatmc.so).  All ranks call; the result is used on rank 0."""
    out = {}
    # ---- cfg 3, strong: 1e5 pairs in total, by pair, counters all-gathered -------------------------------------
    pairs = wl.dataset_pairs(100_000, seed=3)
    n_pairs, n_samples = pairs.size, 10_000
    d_pairs = torch.from_numpy(np.ascontiguousarray(pairs).view(np.float32)).cuda()
    d_hits = torch.zeros(group.hits_capacity(n_pairs), dtype=torch.int64, device="cuda")
    for s in range(3):
        group.count_fused(d_pairs, n_pairs, n_samples, 7 + s, mod.SHARD_BY_PAIR, d_hits)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    for s in range(steps):
        flush.fill_(s & 0xff)
        ev[s][0].record()
        group.count_fused(d_pairs, n_pairs, n_samples, 100 + s, mod.SHARD_BY_PAIR, d_hits)
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    group.set_timing(True)                                     # split one step into kernel and collective (synchronous)
    k_ms, c_ms = [], []
    for s in range(5):
        barrier()
        group.count_fused(d_pairs, n_pairs, n_samples, 200 + s, mod.SHARD_BY_PAIR, d_hits)
        k, c = group.last_times(); k_ms.append(k); c_ms.append(c)
    group.set_timing(False)
    h_pairs, p1 = pinned_array(mod, n_pairs * 48, np.uint8)
    h_hits, p2 = pinned_array(mod, n_pairs * 8, np.uint64)
    h_pairs[:] = np.frombuffer(pairs.tobytes(), dtype=np.uint8)
    hp = h_pairs.view(mod.PAIR_DTYPE)
    for s in range(2):
        group.count_fused_host(hp, n_samples, 7 + s, mod.SHARD_BY_PAIR, out=h_hits)
    barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        group.count_fused_host(hp, n_samples, 300 + s, mod.SHARD_BY_PAIR, out=h_hits)
    barrier()
    e2e = time.perf_counter() - t0
    t = torch.tensor([dev_ms, wall * 1e3, e2e * 1e3, float(np.median(k_ms)), float(np.median(c_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, k_med, c_med = (float(x) for x in t.cpu())
    tests = float(n_pairs) * n_samples
    out["cfg3"] = {"what": "1e5 pairs x 1e4 samples in TOTAL, sharded by pair (ceil(n/N) pairs per rank), counters all-gathered to "
                           "every rank by ncclAllGather inside satmc_group_count_fused; inputs resident",
                   "tests_per_s": tests * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps, "wall_ms_per_step": wall_ms / steps,
                   "steps": steps, "kernel_ms": k_med, "allgather_us": c_med * 1e3,
                   "e2e_tests_per_s": tests * steps / (e2e_ms * 1e-3), "e2e_ms_per_step": e2e_ms / steps,
                   "e2e_api": "satmc_group_count_fused_host: each rank uploads its slice (pinned H2D), counts, all-gathers, downloads all counters",
                   "pairs_per_rank": -(-n_pairs // world), "hits_checksum": int(d_hits[:n_pairs].sum().item())}
    lib = mod.load_library()
    lib.satmc_host_free(p1); lib.satmc_host_free(p2)

    # ---- cfg 4, strong: one pair, 1e11 samples in total, by sample range + ONE all-reduce of the u64 count -------
    one = wl.cfg2_pair()
    d_one = torch.from_numpy(np.ascontiguousarray(one).view(np.float32)).cuda()
    d_h1 = torch.zeros(max(1, group.hits_capacity(1)), dtype=torch.int64, device="cuda")
    total = 100_000_000_000
    group.count_fused(d_one, 1, total // 50, 5, mod.SHARD_BY_SAMPLE_RANGE, d_h1)       # warm-up (2e9 samples)
    barrier()
    n4 = 3
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n4)]
    counts = []
    t0 = time.perf_counter()
    for s in range(n4):
        ev[s][0].record()
        group.count_fused(d_one, 1, total, 1000 + s, mod.SHARD_BY_SAMPLE_RANGE, d_h1)
        ev[s][1].record()
        counts.append(int(d_h1[0].item()))                      # the caller reads the count: D2H of 8 bytes inside the wall time
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    group.set_timing(True)
    barrier()
    group.count_fused(d_one, 1, total, 1000, mod.SHARD_BY_SAMPLE_RANGE, d_h1)
    k_ms, c_ms = group.last_times()
    # the all-reduce alone: the same call on a range so short that the kernels are empty, ranks aligned by a barrier
    lat = []
    for s in range(20):
        barrier()
        group.count_fused(d_one, 1, 4 * world, 9, mod.SHARD_BY_SAMPLE_RANGE, d_h1)
        lat.append(group.last_times()[1])
    group.set_timing(False)
    t = torch.tensor([dev_ms, wall * 1e3, k_ms, c_ms, float(np.median(lat))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, k_ms, c_ms, lat_med = (float(x) for x in t.cpu())
    out["cfg4"] = {"what": "ONE pair, 1e11 samples in TOTAL, sharded by sample range, one ncclAllReduce(ncclUint64, ncclSum) of the "
                           "count inside satmc_group_count_fused (the reference cannot run this: int n_samples, float counter)",
                   "tests_per_s": float(total) * n4 / (dev_ms * 1e-3), "ms_per_step": dev_ms / n4, "wall_ms_per_step_incl_readback": wall_ms / n4,
                   "steps": n4, "samples_per_rank": total // world, "kernel_ms": k_ms,
                   "allreduce_us": c_ms * 1e3, "allreduce_us_note": "time between the end of this rank's kernel and the end of the "
                   "collective in one full step (includes waiting for the slowest rank), max over ranks",
                   "allreduce_only_us": lat_med * 1e3, "allreduce_only_note": "median over 20 calls with empty kernels after a barrier: "
                   "the latency of the 8-byte all-reduce itself" if world > 1 else "world size 1: no collective is issued",
                   "counts": counts, "p_hat": counts[0] / float(total)}

    # ---- sharding check: the N-rank count of a 1e9-sample range == one GPU's count of the same range ------------
    n_chk, off = 1_000_000_000, 123_456_789_012
    group.count_fused(d_one, 1, n_chk, 4, mod.SHARD_BY_SAMPLE_RANGE, d_h1, sample_offset=off)
    torch.cuda.synchronize()
    got = int(d_h1[0].item())
    d_ref = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.count_fused(d_one, 1, n_chk, 4, d_ref, sample_offset=off)                     # every rank alone, whole range
    torch.cuda.synchronize()
    want = int(d_ref[0].item())
    ok = torch.tensor([1.0 if got == want else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res = {"samples": n_chk, "sample_offset": off, "n_rank_count": got, "single_gpu_count": want,
           "bit_equal_on_every_rank": bool(ok.item() == 1.0), "ranks": world}
    if not res["bit_equal_on_every_rank"]:
        raise SystemExit(f"sharding check FAILED: {res}")
    return {"strong": out, "sharding_check": res}


def programs_gpu_count_invariance(dist, rank, world):
    """Rank 0 runs the drop-in programs as a user would: generate_dataset once, then ztest on the same rows with --gpus 1 and
    --gpus N (single process, satmc_group_create inside): the outputs must be identical.  The other ranks wait on the store
    (no NCCL kernel spinning on their GPUs meanwhile)."""
    store = dist.distributed_c10d._get_default_store()
    res = None
    if rank == 0:
        host = os.path.join(ROOT, PKG, "host")
        d = tempfile.mkdtemp(prefix="satmc_bench_")
        try:
            def run(cmd):
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                if r.returncode != 0:
                    raise RuntimeError(f"{' '.join(cmd)} -> {r.returncode}: {r.stderr[-400:]}")
                return time.perf_counter() - t0
            t_gen = run([os.path.join(host, "generate_dataset"), "--data_dir", d, "-n", "1", "-b", "200000", "--num_poses", "65536",
                         "--num_variances", "65536", "--max_samples", "400000", "--seed", "5"])
            data = np.load(os.path.join(d, "0.npy"))
            os.makedirs(os.path.join(d, "tmp"), exist_ok=True)
            np.save(os.path.join(d, "tmp", "0.npy"), np.ascontiguousarray(data[:, [0, 1, 3, 4]], dtype=np.float32))
            outs, times = {}, {}
            for g in (1, world):
                f = os.path.join(d, f"out{g}.npy")
                times[g] = run([os.path.join(host, "ztest"), "--data_dir", d, "--data_file_out", f, "--max_samples", "400000",
                                "--seed", "7", "--gpus", str(g), "--meta_dir", os.path.join(d, "meta")])
                outs[g] = np.load(f)
            same = bool(np.array_equal(outs[1], outs[world]))
            res = {"rows": int(data.shape[0]), "generate_dataset_wall_s": t_gen, "ztest_gpus_1_wall_s": times[1],
                   f"ztest_gpus_{world}_wall_s": times[world], "outputs_identical": same, "mean_cp": float(outs[1][:, 2].mean()),
                   "what": "ztest --gpus N deals the rows round-robin to N GPUs (satmc_group_adaptive_run_host); wall times include "
                           "process start-up and CUDA/NCCL initialisation"}
            if not same:
                raise RuntimeError("ztest output depends on the number of GPUs")
            res["single_process_group"] = single_process_group(world)
        except Exception as e:                                # reported, not fatal for the bench line
            res = {"error": str(e)[:400]}
        finally:
            shutil.rmtree(d, ignore_errors=True)
            store.set("satmc_programs_done", "1")
    else:
        store.wait(["satmc_programs_done"])
    return res


def single_process_group(world):
    """ONE process driving all N GPUs (satmc_group_create), cfg 4 through host buffers: sample ranges combined either by the
    counting kernels' own epilogue (system-scope atomics into device 0's memory over NVLink, no collective launch) or by
    ncclAllReduce.  Wall clock per call, host buffers in and out; a 1e7-sample call exposes the fixed cost of each path."""
    mod = importlib.import_module(PKG)
    wl = importlib.import_module(PKG + ".workloads")
    one = wl.cfg2_pair()
    out = {}
    with mod.Group(devices=list(range(world))) as g:
        ref = None
        for label, on in (("peer_atomics", True), ("nccl", False)):
            g.set_peer_reduce(on)
            g.count_fused_host(one, 2_000_000_000, 3, mod.SHARD_BY_SAMPLE_RANGE)
            t0 = time.perf_counter()
            for s_ in range(2):
                big = g.count_fused_host(one, 100_000_000_000, 1000 + s_, mod.SHARD_BY_SAMPLE_RANGE)
            t_big = (time.perf_counter() - t0) / 2
            ts = []
            for s_ in range(60):
                t0 = time.perf_counter()
                small = g.count_fused_host(one, 10_000_000, 7, mod.SHARD_BY_SAMPLE_RANGE)
                ts.append(time.perf_counter() - t0)
            out[label] = {"exchange": g.last_exchange(), "cfg4_1e11_ms_per_call": t_big * 1e3, "cfg4_1e11_tests_per_s": 1e11 / t_big,
                          "call_1e7_samples_us": float(np.median(ts[10:])) * 1e6, "count_1e11": int(big[0]), "count_1e7": int(small[0])}
            ref = ref or (int(big[0]), int(small[0]))
            out[label]["same_counts_as_first_path"] = bool((int(big[0]), int(small[0])) == ref)
    out["what"] = "satmc_group_count_fused_host on N GPUs from one process: wall clock incl. H2D of the pair and D2H of the count"
    return out


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
    base = wl.dataset_pairs(1000, 5); d_base = put(base)
    ms = timed(lambda: ctx.count_fused_sweep(d_base, base.size, sig, 64, 100_000, 7, d_hs), reps=3)
    out["cfg5_fused_sweep_common_random_numbers"] = {"tests_per_s": base.size * 64 * 1e5 / ms * 1e3, "ms": ms, "rows": int(base.size * 64),
                                                     "samples": 100_000, "what": "satmc_count_fused_sweep: 64 covariance settings per pair "
                                                     "share one Philox stream (the sampler runs once per sample, not once per setting)"}
    zb = torch.randn(3 * 100_000, device="cuda")
    ms = timed(lambda: ctx.count_streamed(d_sw, sweep.size, zb, 100_000, 3, 100_000, d_hs), reps=3)
    out["cfg5_streamed_shared_bank"] = {"tests_per_s": sweep.size * 1e5 / ms * 1e3, "ms": ms}
    out["cfg2"] = cfg2_latency(mod, wl, torch)
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
        ctx.exact_evals(reset=True)
        ms = timed(lambda: ctx.count_fused_polygons(d_poly, ppoly.size, 10_000, 7, d_hp), reps=3, warm=0)
        out[name] = {"tests_per_s": ppoly.size * 1e4 / ms * 1e3, "ms": ms, "mean_p": float(d_hp.sum().item()) / (ppoly.size * 1e4),
                     "exact_pass_fraction": ctx.exact_evals(reset=True) / (3.0 * ppoly.size * 1e4),
                     "what": "satmc_count_fused_polygons: screening pass, undecided samples compacted per warp, exact polygon SAT"}
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
    l0 = ctx.launch_count
    dts = []
    for _ in range(3):                                               # wall clock: the median of three
        t0 = time.perf_counter(); iters, drawn = adaptive(); torch.cuda.synchronize(); dts.append(time.perf_counter() - t0)
    dt = sorted(dts)[1]
    out["adaptive_batch"] = {"pairs": int(pairs.size), "max_samples": max_samples, "ms": dt * 1e3, "iterations": iters,
                             "ms_runs": [round(t * 1e3, 3) for t in dts], "kernel_launches": (ctx.launch_count - l0) // 3,
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


def cfg2_latency(mod, wl, torch):
    """BASELINE config 2 (one pair x 1e6 samples) is a latency measurement: 3.3 us of work at the steady rate.  Reported: device
    time per call (CUDA events around single calls), wall time per call for a caller that needs every result (launch +
    synchronise), wall time per call for back-to-back asynchronous calls, and the same as a replayed CUDA graph."""
    s = torch.cuda.Stream()
    out = {}
    with torch.cuda.stream(s):
        c2 = mod.Context(0, s.cuda_stream)
        one = torch.from_numpy(np.ascontiguousarray(wl.cfg2_pair()).view(np.float32)).cuda()
        d_h = torch.zeros(1, dtype=torch.int64, device="cuda")
        for _ in range(20):
            c2.count_fused(one, 1, 1_000_000, 7, d_h)
        s.synchronize()
        l0 = c2.launch_count
        c2.count_fused(one, 1, 1_000_000, 7, d_h)
        launches_per_call = c2.launch_count - l0
        ts = []
        for _ in range(200):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); c2.count_fused(one, 1, 1_000_000, 7, d_h); b.record(s); s.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        out["device_us_per_call"] = float(np.median(ts))
        t0 = time.perf_counter()
        for _ in range(500):
            c2.count_fused(one, 1, 1_000_000, 7, d_h); s.synchronize()
        out["wall_us_per_call_launch_and_sync"] = (time.perf_counter() - t0) / 500 * 1e6
        t0 = time.perf_counter()
        for _ in range(2000):
            c2.count_fused(one, 1, 1_000_000, 7, d_h)
        s.synchronize()
        out["wall_us_per_call_back_to_back"] = (time.perf_counter() - t0) / 2000 * 1e6
        want = int(d_h.item())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                c2.count_fused(one, 1, 1_000_000, 7, d_h)
        g.replay(); s.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            g.replay()
        s.synchronize()
        out["wall_us_per_call_cuda_graph_of_10"] = (time.perf_counter() - t0) / 2000 * 1e6
        out["graph_result_equal"] = bool(int(d_h.item()) == want)
        out["kernel_launches_per_call"] = int(launches_per_call)
        out["tests_per_s_device"] = 1e6 / (out["device_us_per_call"] * 1e-6)
        out["hits"] = want
        c2.close()
    return out


if __name__ == "__main__":
    main()
