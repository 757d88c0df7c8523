#!/bin/bash
# build_variant.sh NAME "-DFLAG=.." : builds build/libfod_NAME.so with extra nvcc flags (kernel A/B experiments)
set -e
cd "$(dirname "$0")/../faster_orefsdet_b200/csrc"
mkdir -p ../../build/$1
for f in api tmap nms decode correlate correlate_tc roi relation_tc conv_tc stem1_tc gn glue; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --fmad=true $2 -c $f.cu -o ../../build/$1/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libfod_$1.so ../../build/$1/*.o -lcudart
