#!/bin/bash
# Builds libpixelnerf_b200.so in-tree for sm_100a (the only supported target).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${PNR_OUT:-$HERE/../libpixelnerf_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr ${PNR_EXTRA_NVCC_FLAGS}"
BUILD="${PNR_BUILD_DIR:-$HERE/build}"
mkdir -p "$BUILD"
pids=()
for f in api features mlp_f32 rays mlp_tc tc_probe; do
  ( $NVCC $FLAGS -c "$HERE/$f.cu" -o "$BUILD/$f.o" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$BUILD"/{api,features,mlp_f32,rays,mlp_tc}.o -lcudart -ldl
# hardware probes (tests/test_gpu_tc_probe.py, tools/probe_*.py): their own library, not part of the product ABI
OUTDIR="$(cd "$(dirname "$OUT")" && pwd)"
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUTDIR/libpixelnerf_b200_probe.so" "$BUILD/tc_probe.o" \
  -L"$OUTDIR" -l:"$(basename "$OUT")" -Xlinker -rpath -Xlinker '$ORIGIN' -lcudart
echo "built $OUT"
