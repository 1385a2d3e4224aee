#!/bin/bash
# three-group kernel (engine 6) with the hand-pipelined exponential section: kernel A/B and tests
mkdir -p gpurun_out
OUT=gpurun_out/r2g_attn3_swp.txt; : > $OUT
for lib in build/variants/libpd_t3s*.so; do
  echo "== $lib" >> $OUT
  PD_B200_LIB=$PWD/$lib timeout 200 python scripts/attn_bench.py 2>&1 | grep -E " 6 \|" >> $OUT
done
cat $OUT
PD_B200_LIB=$PWD/build/variants/libpd_t3s2.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -2
