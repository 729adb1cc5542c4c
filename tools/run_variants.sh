# development helper: adaptive batch, current planner against the forced resident cut and an older build
echo "== shipped"; python tools/probe_adaptive.py 2>&1 | tail -1
echo "== SATMC_TINY_BPS=2"; SATMC_TINY_BPS=2 python tools/probe_adaptive.py 2>&1 | tail -1
echo "== build 3ba15c1"; SATMC_LIB=$PWD/variants/c3ba/libsatmc.so LD_LIBRARY_PATH=$PWD/variants/c3ba python tools/probe_adaptive.py 2>&1 | tail -1
echo "== shipped again"; python tools/probe_adaptive.py 2>&1 | tail -1
