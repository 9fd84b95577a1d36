#!/bin/bash
# usage: run.sh variant cases...
cd /root/repo
for v in "$@"; do
  cp tmp_variants/lib_$v.so modcholesky_b200/csrc/libmodchol_b200.so
  echo "VARIANT $v"
  timeout 300 python tools/quick_bench.py --cases se99:32:1048576,gmw81:32:1048576 --families uniform,wishart
done
