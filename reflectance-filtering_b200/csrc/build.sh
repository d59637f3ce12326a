#!/bin/bash
# Builds librf_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -Xcompiler -Wall -Wno-deprecated-gpu-targets -ccbin /usr/bin/g++)
OBJS=()
PIDS=()
for f in misc bf bf2 bf_generic cnn cnn_tc gf gf2 colorize whdr; do
  rm -f "$f.o"
  "$NVCC" "${FLAGS[@]}" ${RF_PTXAS_V:+-Xptxas -v} -c "$f.cu" -o "$f.o" &
  PIDS+=($!)
  OBJS+=("$f.o")
done
for p in "${PIDS[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o librf_b200.so "${OBJS[@]}" -ccbin /usr/bin/g++
echo "built $(pwd)/librf_b200.so"
