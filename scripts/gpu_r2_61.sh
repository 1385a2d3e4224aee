#!/bin/bash
# ptxas optimisation level of the GEMM engine's translation unit: same-box A/B of the graph-replayed step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2g_ab_gemm_ptxas.txt; : > $OUT
for i in 1 2; do
  for lib in build/variants/libpd_g3.so build/variants/libpd_g2.so build/variants/libpd_g1.so; do
    echo -n "$lib: " >> $OUT
    PD_B200_LIB=$PWD/$lib timeout 300 python scripts/profile_step.py --graph 1 --reps 20 2>&1 | tail -1 >> $OUT
  done
done
cat $OUT
