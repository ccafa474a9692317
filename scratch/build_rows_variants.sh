#!/bin/bash
# builds libse3et_b200.so variants that differ in -D flags of kpconv_rows.cu: scratch/variants/<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/variants /tmp/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DSE3ET_ROWS_DEV"
names=""
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  nvcc $FLAGS $defs -c se3et_b200/csrc/kpconv_rows.cu -o /tmp/variants/${name}_rows.o &
  names="$names $name"
done
wait
for name in $names; do
  objs=""
  for f in se3et_b200/csrc/*.cu; do
    b=$(basename $f .cu)
    if [ $b = kpconv_rows ]; then objs="$objs /tmp/variants/${name}_rows.o"; else objs="$objs se3et_b200/csrc/$b.o"; fi
  done
  nvcc -shared -Wno-deprecated-gpu-targets -o scratch/variants/$name.so $objs -lcudart
  echo "built scratch/variants/$name.so"
done
