#!/usr/bin/env bash
# oracle/build_ref.sh -- compile the UNMODIFIED reference (thvasilo/cuda-sgd-sese-project) from the
# sources where they lie under /root/reference/c_code into oracle/_ref/ (git-ignored, but shipped
# to the GPU box by gpurun).  No reference source is copied into the repository: the translation
# units are compiled in place; object files and the binaries are the only outputs.
#
# The reference is a CUDA program (Thrust + cuBLAS): it has NO CPU path and can only RUN on the GPU
# box.  There it serves (a) as the second pin of the oracle (tests/test_reference_binary.py) and
# (b) as the "reference cuBLAS codepath on one B200" baseline of bench.py --impl reference.
#
# Deviations from c_code/Makefile forced by this image:
#   -std=c++17 (CCCL 2.8 requires it), -arch=sm_100a, -lcublas instead of the static/device cuBLAS
#   of CUDA 7.5, and an EMPTY sampling.cuh: main.cu:4 includes a header that is not in the
#   reference repository.  The stub is generated into the build directory, never committed.
#
# Two binaries are produced:
#   _ref/main          the reference exactly as it is;
#   _ref/main_weights  the same, except that the commented-out `print_vector(weights, ...)`
#                      (main.cu:560) is enabled by a sed pass over a temporary copy of main.cu in
#                      $TMPDIR, so that the final weights can be compared (three-way parity).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}/c_code"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) -- keeping any prebuilt $OUT" >&2
  exit 0
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -lineinfo -w"
mkdir -p "$OUT/obj"
STUB="$(mktemp -d)"
trap 'rm -rf "$STUB"' EXIT
: > "$STUB/sampling.cuh"
for f in jsoncpp.cpp sgd_io.cu sgd_thrust.cu testing.cu main.cu; do
  o="$OUT/obj/${f%.*}.o"
  if [ ! -f "$o" ] || [ "$REF/$f" -nt "$o" ]; then
    "$NVCC" $FLAGS -I"$REF" -I"$STUB" -x cu -dc "$REF/$f" -o "$o"
  fi
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a "$OUT"/obj/{jsoncpp,sgd_io,sgd_thrust,testing,main}.o \
  -lcudadevrt -lcublas -o "$OUT/main"
# weights-dumping variant: temporary patched copy of main.cu, outside the repository
sed 's|^\(\s*\)// print_vector(weights, "weights");|\1std::cout.precision(9); print_vector(weights, "weights");|' \
  "$REF/main.cu" > "$STUB/main_weights.cu"
grep -q 'std::cout.precision(9); print_vector(weights' "$STUB/main_weights.cu"
"$NVCC" $FLAGS -I"$REF" -I"$STUB" -x cu -dc "$STUB/main_weights.cu" -o "$OUT/obj/main_weights.o"
"$NVCC" -gencode arch=compute_100a,code=sm_100a "$OUT"/obj/{jsoncpp,sgd_io,sgd_thrust,testing,main_weights}.o \
  -lcudadevrt -lcublas -o "$OUT/main_weights"
rm -rf "$OUT/obj"
echo "built $OUT/main and $OUT/main_weights"
