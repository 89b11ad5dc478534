#!/usr/bin/env bash
# usage: tools/run_variants.sh var1.so var2.so ...   (files under dynlab_b200/_lib)
for v in "$@"; do
  echo "== $v"
  DLB_LIB_PATH=$PWD/dynlab_b200/_lib/$v python tools/time_k1.py 2>&1 | grep -E "K1|K2"
done
