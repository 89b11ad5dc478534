#!/usr/bin/env bash
for v in "$@"; do
  echo "== $v"
  DLB_LIB_PATH=$PWD/dynlab_b200/_lib/$v python tools/time_k3.py 2>&1 | grep -E "s_min  |rate"
  DLB_LIB_PATH=$PWD/dynlab_b200/_lib/$v python tools/time_k1.py 2>&1 | grep -E "K2"
done
