run() { # name lib chunk
  echo "== $1 chunk=$3"
  if [ "$2" = shipped ]; then SATMC_STREAM_CHUNK=$3 python tools/quick_bench.py --only --streamed 2>&1 | grep -E "private|rror"
  else SATMC_STREAM_CHUNK=$3 SATMC_LIB=$PWD/variants/$2/libsatmc.so LD_LIBRARY_PATH=$PWD/variants/$2 python tools/quick_bench.py --only --streamed 2>&1 | grep -E "private|rror"; fi
}
for v in shipped s4t128 s2t256 s4t256; do for c in 0 8192 4096 2048; do run $v $v $c; done; done
run nofinal nofinal 0
run unpacked unpacked 0
