#!/bin/bash
for v in pipe96 pipe112; do
  cp gpurun_out/lib_$v.so tinyrecurrentunet_b200/libtru_b200.so
  echo "== $v"; timeout 120 python scratch/bwd_tc.py 2>&1 | grep -A1 "shape\|full" | grep "shape\|full"
done
