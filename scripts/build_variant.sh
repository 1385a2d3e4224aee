#!/bin/bash
# Build an experimental copy of the library with extra -D flags for ONE translation unit:
#   scripts/build_variant.sh <tag> <file-stem> "-DFOO=1 -DBAR=2"   ->  build/variants/libpd_<tag>.so
# (select it with PD_B200_LIB=build/variants/libpd_<tag>.so; the other objects come from the last ./build.sh)
set -e
cd "$(dirname "$0")/.."
tag=$1; stem=$2; defs=$3
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
if [ "$stem" == "attention_tc" ] && [[ "$defs" != *"-Xptxas"* ]]; then defs="$defs -Xptxas -O1"; fi   # as in build.sh
nvcc $FLAGS $defs -c prompt-diffusion_b200/csrc/$stem.cu -o build/variants/${stem}_$tag.o 2>&1 | grep -v "deprecated-gpu-targets" || true
objs=""
for f in elementwise norm softmax conv_simt attention_simt attention_mma attention_tc attention_tc3 attention_tc4 attention_xtc attention_ptc attention_short gemm_sm100; do
  if [ "$f" == "$stem" ]; then objs="$objs build/variants/${stem}_$tag.o"; else objs="$objs build/$f.o"; fi
done
nvcc -shared -o build/variants/libpd_$tag.so $objs -lcudart_static -ldl -lrt -lpthread 2>&1 | grep -v "deprecated-gpu-targets" || true
echo "built build/variants/libpd_$tag.so"
