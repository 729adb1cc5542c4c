# development helper: time the streamed kernels of several library builds inside one gpurun call (same box)
run() { # lib
  echo "== $1 shape=${QB_NPAIRS:-16384}x${QB_N:-32768}"
  if [ "$1" = shipped ]; then python tools/quick_bench.py --only --streamed 2>&1 | grep -E "streamed|rror"
  else SATMC_LIB=$PWD/variants/$1/libsatmc.so LD_LIBRARY_PATH=$PWD/variants/$1 python tools/quick_bench.py --only --streamed 2>&1 | grep -E "streamed|rror"; fi
}
for v in c3ba shipped; do run $v; done
export QB_NPAIRS=16001 QB_N=33920
run shipped
