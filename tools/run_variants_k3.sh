#!/usr/bin/env bash
# usage: [VENV="K=V ..."] tools/run_variants_k3.sh var1.so var2.so ...   (files under dynlab_b200/_lib)
for v in "$@"; do
  echo "== $v"
  DLB_LIB_PATH=$PWD/dynlab_b200/_lib/$v timeout -s KILL 120 python tools/time_k3.py 2>&1 | grep -E "s_min|rate "
  DLB_LIB_PATH=$PWD/dynlab_b200/_lib/$v timeout -s KILL 120 python tools/time_k1.py 2>&1 | grep -E "K2"
done
