#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a (B200) in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../_lib"
mkdir -p "$OUT" "$HERE/_obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
       -Xcompiler -fPIC -Xptxas -v ${DLB_NVCC_EXTRA:-})
pids=()
for f in capi rk45 stencil floweval contour select link eulerflow peer; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" > "$HERE/_obj/$f.log" 2>&1 \
      || { cat "$HERE/_obj/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libdynlab_b200.so" "$HERE"/_obj/{capi,rk45,stencil,floweval,contour,select,link,eulerflow,peer}.o
echo "built $OUT/libdynlab_b200.so"
