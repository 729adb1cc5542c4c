#!/usr/bin/env bash
# oracle/build_ref.sh -- compile the UNMODIFIED reference for sm_100a into oracle/_ref/.
#
# Test infrastructure only.  The sources are read where they lie under $REF (default
# /root/reference); nothing is copied.  The reference needs boost headers and libnpy, which
# this image lacks: oracle/stubs/ holds compile-only stand-ins (the harness never calls the
# reference's main/parse_args/file I/O).  thrust/count.h and thrust/sort.h are force-included
# because the reference relies on transitive includes that CCCL 2.8 no longer provides.
# Default nvcc floating-point flags (-fmad=true, no fast-math) -- the same contraction the
# reference's own "nvcc -o ..." build line (README.md:6-7) gets.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/ztest.cu" ]; then
    echo "build_ref: $REF/ztest.cu not found -- keeping prebuilt $OUT (GPU box case)" >&2
    exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -I "$HERE/stubs" -I "$REF" -include thrust/count.h -include thrust/sort.h \
    -Xcompiler -fPIC -shared -w -o "$OUT/libref_gpu.so" "$HERE/ref_gpu.cu"
echo "build_ref: built $OUT/libref_gpu.so"
