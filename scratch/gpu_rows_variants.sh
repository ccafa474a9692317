#!/bin/bash
cp se3et_b200/csrc/libse3et_b200.so /tmp/orig.so
for v in "$@"; do
  cp scratch/variants/$v.so se3et_b200/csrc/libse3et_b200.so
  touch se3et_b200/csrc/.build_stamp
  echo "== $v"
  SE3ET_NO_REBUILD=1 timeout 300 python scratch/bench_rows.py 16 2>&1 | tail -6
done
cp /tmp/orig.so se3et_b200/csrc/libse3et_b200.so
