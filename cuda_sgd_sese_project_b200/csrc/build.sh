#!/usr/bin/env bash
# Builds libb200sgd.so (the C-ABI product library) and the `main` CLI for sm_100a, in-tree.
# nvcc cross-compiles without a GPU.  Usage: bash build.sh [-j]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/.."
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
CXXFLAGS="-O3 -std=c++17 -lineinfo -I$HERE -I$HERE/../../include"
mkdir -p "$HERE/obj"
g++ -O3 -std=c++17 -fPIC -Wall -Wextra -I"$HERE" -c "$HERE/host_io.cpp" -o "$HERE/obj/host_io.o" &
"$NVCC" $ARCH $CXXFLAGS -Xcompiler -fPIC,-Wall -Xptxas -v -c "$HERE/b200sgd_api.cu" -o "$HERE/obj/b200sgd_api.o" \
  2> "$HERE/obj/ptxas.log" &
wait
"$NVCC" $ARCH -shared -o "$OUT/libb200sgd.so" "$HERE/obj/b200sgd_api.o" "$HERE/obj/host_io.o" -lpthread
g++ -O2 -std=c++17 -Wall -Wextra -I"$HERE/../../include" "$HERE/main.cpp" -o "$OUT/main" \
  -L"$OUT" -lb200sgd -Wl,-rpath,'$ORIGIN'
echo "built $OUT/libb200sgd.so and $OUT/main"
