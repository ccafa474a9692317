#!/bin/bash
# builds libse3et_b200.so variants that differ in -D flags of the tcgen05 kernels: scratch/variants/<name>.so
# usage: build_variants.sh name1 "-DX=1 -DY=2" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
python -c "from se3et_b200 import build; build.build_library()"   # current objects for everything else
mkdir -p scratch/variants /tmp/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  objs=""
  for f in se3et_b200/csrc/*.cu; do
    b=$(basename $f .cu)
    case $b in
      kpconv_fused|gemm|transformer)
        nvcc $FLAGS $defs -c $f -o /tmp/variants/${name}_$b.o &
        objs="$objs /tmp/variants/${name}_$b.o";;
      *) objs="$objs se3et_b200/csrc/$b.o";;
    esac
  done
  wait
  nvcc -shared -Wno-deprecated-gpu-targets -o scratch/variants/$name.so $objs -lcudart
  echo "built scratch/variants/$name.so"
done
