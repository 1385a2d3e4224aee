#!/bin/bash
# Build the C-ABI library (sm_100a only) in-tree: prompt-diffusion_b200/libpd_b200.so
set -e
cd "$(dirname "$0")/prompt-diffusion_b200/csrc"
OUT=../libpd_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math"
FLAGS_ACC="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p ../../build
pids=()
for f in elementwise norm softmax conv_simt attention_simt; do
  nvcc $FLAGS_ACC -c $f.cu -o ../../build/$f.o & pids+=($!)
done
for f in attention_mma attention_tc3 attention_tc4 attention_xtc attention_ptc attention_short gemm_sm100; do
  nvcc $FLAGS_ACC -c $f.cu -o ../../build/$f.o & pids+=($!)
done
# attention_tc.cu carries a HAND-SCHEDULED exponential section (FA_HAND): ptxas -O1 keeps the interleave of SFU and
# FMA-pipe work written in the source, the default -O3 hoists the polynomial chains in front of the SFU work
# (measured 639 vs 712 us at B16 h8 N4096 d40, profiles/r02_attn_hand.txt)
nvcc $FLAGS_ACC -Xptxas -O1 -c attention_tc.cu -o ../../build/attention_tc.o & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT.tmp ../../build/elementwise.o ../../build/norm.o ../../build/softmax.o ../../build/conv_simt.o \
  ../../build/attention_simt.o ../../build/attention_mma.o ../../build/attention_tc.o ../../build/attention_tc3.o ../../build/attention_tc4.o ../../build/attention_xtc.o ../../build/attention_ptc.o ../../build/attention_short.o ../../build/gemm_sm100.o -lcudart_static -ldl -lrt -lpthread
mv -f $OUT.tmp $OUT
echo "built $(realpath $OUT)"
