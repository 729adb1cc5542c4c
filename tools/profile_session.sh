#!/bin/bash
# tools/profile_session.sh -- one GPU box: everything profiles/ keeps for a round (run through gpurun; outputs under gpurun_out/).
# usage: tools/profile_session.sh <tag>        e.g. r2
set -u
T=${1:-r2}
P=convex-2d-gpu-collision-detection_b200
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > $O/${T}_gpu.txt

# 1. GPU test suite: shipped library, then the -DSATMC_DEBUG build (device-side bounds asserts)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${T}_gpu_suite.log
( SATMC_LIB=$PWD/$P/debug/libsatmc.so LD_LIBRARY_PATH=$PWD/$P/debug python -c "import importlib; m = importlib.import_module('$P'); print('library under test:', m.load_library().satmc_version().decode(), m.LIB_PATH)";
  SATMC_LIB=$PWD/$P/debug/libsatmc.so LD_LIBRARY_PATH=$PWD/$P/debug python -m pytest tests -m gpu -q 2>&1 ) > $O/${T}_debug_full.log
( head -8 $O/${T}_debug_full.log; echo "..."; tail -6 $O/${T}_debug_full.log ) > $O/${T}_debug_suite.log; rm -f $O/${T}_debug_full.log

# 2. bench lines
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench_n1.err

# 3. launch lists (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-strong > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_adaptive.csv \
    python tools/prof_fused.py adaptive > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $O/${T}_launches_cfg2.csv \
    python tools/prof_fused.py cfg2 > /dev/null 2>&1

# 4. full captures of the counting kernels (one launch each, after the same command ran clean above)
# (the reports are condensed to text on the box: gpurun merges at most 64 MiB back; only the fused one travels whole)
for m in fused fused5 cfg4 cfg2 sweep poly streamed3 streamed5; do
    ncu --set full --clock-control none --import-source on -k regex:k_count --launch-skip 1 --launch-count 1 -f -o /tmp/${T}_ncu_$m \
        python tools/prof_fused.py $m > /dev/null 2>&1
    python tools/ncu_summary.py --json=$O/${T}_ncu_$m.json /tmp/${T}_ncu_$m.ncu-rep > $O/${T}_ncu_$m.txt 2>&1
done
cp /tmp/${T}_ncu_fused.ncu-rep $O/
ncu --set full --clock-control none --import-source on -k regex:k_count --launch-skip 8 --launch-count 1 -f -o /tmp/${T}_ncu_indirect \
    python tools/prof_fused.py adaptive > /dev/null 2>&1
python tools/ncu_summary.py /tmp/${T}_ncu_indirect.ncu-rep > $O/${T}_ncu_indirect.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ztest_compact --launch-skip 2 --launch-count 1 -f -o /tmp/${T}_ncu_ztest_compact \
    python tools/prof_fused.py adaptive > /dev/null 2>&1
python tools/ncu_summary.py /tmp/${T}_ncu_ztest_compact.ncu-rep > $O/${T}_ncu_ztest_compact.txt 2>&1

# 5. the drop-in programs: where the time goes
rm -rf /tmp/gd && ( time $P/host/generate_dataset --data_dir /tmp/gd -n 100 --seed 1 --stats ) 2>&1 | tr '\r' '\n' | grep -E 'stats|real' > $O/${T}_programs.txt
python - <<'PY' >> $O/${T}_programs.txt 2>&1
import numpy as np, os, shutil
d = "/tmp/gd"; din = "/tmp/gd_in"; dout = "/tmp/gd_out"
shutil.rmtree(din, ignore_errors=True); shutil.rmtree(dout, ignore_errors=True); os.makedirs(din); os.makedirs(dout + "/meta")
for b in range(20):
    a = np.load(f"{d}/{b}.npy"); np.save(f"{din}/{b}.npy", np.ascontiguousarray(a[:, [0, 1, 3, 4]], dtype=np.float32))
for f in ("poses.npy", "variances.npy"): shutil.copy(f"{d}/{f}", f"{dout}/{f}")
for f in ("accuracy_bins.npy", "bin_accuracy.npy"): shutil.copy(f"{d}/meta/{f}", f"{dout}/meta/{f}")
print("compute_collision_probability input: 20 files x 100000 rows")
PY
( time $P/host/compute_collision_probability --data_in /tmp/gd_in --data_out /tmp/gd_out --seed 3 --stats ) 2>&1 | tr '\r' '\n' | grep -E 'stats|real' >> $O/${T}_programs.txt
rm -rf /tmp/gd /tmp/gd_in /tmp/gd_out

# 6. pipe-rate microbenchmarks (packed FP32 and the ALU-pipe instructions of the screening loops)
if [ -x tools/ubench ]; then tools/ubench a b c > $O/${T}_ubench_packed.log 2>&1; fi
ls -la $O | tail -40
