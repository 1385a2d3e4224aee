#!/bin/bash
# finer GEMM timelines (CTA 0) of the narrow short-K layer: tabled tile, resident B, pair
mkdir -p gpurun_out
export PD_B200_LIB=build/variants/libpd_dbg.so
{
for shp in "16 64 64 320 320 1 0 1 160 0" "16 64 64 320 320 1 0 1 160 1" "16 64 64 320 320 1 1 1 160 0" "16 64 64 320 320 1 1 1 160 1" "16 64 64 320 320 1 1 2 160 1" "16 32 32 640 640 1 1 1 224 0" "16 16 16 1280 1280 1 1 1 160 0"; do
  timeout 120 python scripts/gemm_timeline.py $shp 2>&1 | head -12
done
} > gpurun_out/r2_12_gemm_timelines.txt
cat gpurun_out/r2_12_gemm_timelines.txt
