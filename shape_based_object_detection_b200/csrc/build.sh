#!/usr/bin/env bash
# Build libsbod.so (sm_100a) in-tree. Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("$here"/*.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared \
  "$@" -o "$out/libsbod.so" "${srcs[@]}"
echo "built $out/libsbod.so"
