#!/usr/bin/env bash
# usage: tools/build_stencil_variant.sh <name> <extra nvcc flags...>
# Recompiles csrc/stencil.cu only and links dynlab_b200/_lib/<name>.so from the existing objects
# (run dynlab_b200/csrc/build.sh first); for A/B runs through DLB_LIB_PATH (tools/run_variants*.sh).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../dynlab_b200/csrc" && pwd)"
name="$1"; shift
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
  -c "$HERE/stencil.cu" -o "$HERE/_obj/stencil_$name.o"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/../_lib/$name.so" \
  "$HERE"/_obj/{capi,rk45,floweval,contour,select,link,eulerflow,peer}.o "$HERE/_obj/stencil_$name.o"
echo "built $name.so"
