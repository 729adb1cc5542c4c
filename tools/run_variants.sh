# development helper: item-length scan of the streamed bulk-tensor kernels (same box, one gpurun call)
for c3 in 1024 1536 2048 2560 3072; do
  echo "== chunk3=$c3"; SATMC_STREAM_CHUNK3=$c3 python tools/quick_bench.py --only --streamed 2>&1 | grep -E "private ndof=3|rror"
done
for c5 in 2048 4096 6144 8192 12288 16384; do
  echo "== chunk5=$c5"; SATMC_STREAM_CHUNK5=$c5 python tools/quick_bench.py --only --streamed 2>&1 | grep -E "private ndof=5|rror"
done
