#!/bin/bash
# Developer tool: A/B builds of libb200gs.so with different -D tuning constants.
#   tools/build_variants.sh name1 "-DX=1 -DY=2" name2 "-D..." ...
# Each variant lands in gpurun_out/variants/libb200gs_<name>.so (select with B200GS_LIB=<path>).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/sdp-gs_b200/csrc
OUT=$ROOT/variants
mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  obj=$OUT/_obj_$name; mkdir -p $obj
  for f in api preprocess binning blend train_ops collective; do
    nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC -Xptxas -v $defs -c $SRC/$f.cu -o $obj/$f.o 2> $obj/$f.log &
  done
  wait
  nvcc -shared $ARCH -o $OUT/libb200gs_$name.so $obj/*.o
  grep -E "Used" $obj/blend.log | sed "s/^/  [$name] blend: /"
done
