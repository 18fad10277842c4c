#!/usr/bin/env bash
# Build libsypha_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
SRC=sypha_b200/csrc
OUT=sypha_b200/lib
mkdir -p "$OUT" build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename "${f%.cu}").o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o")" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libsypha_b200.so" build/*.o
echo "built $OUT/libsypha_b200.so"
