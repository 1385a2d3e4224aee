#!/bin/bash
# stream-K timelines at the 8x8 level (M = 1024): single-CTA 128-wide vs pair tiles
mkdir -p gpurun_out
export PD_B200_LIB=build/variants/libpd_dbg.so
{
for shp in "16 8 8 1280 1280 3 1 1 128 0 1" "16 8 8 1280 1280 3 1 2 128 0 1" "16 8 8 1280 1280 3 1 2 256 0 1" "16 8 8 1280 1280 3 1 1 256 0 1" "16 8 8 2560 1280 3 0 2 256 0 1" "16 16 16 1280 1280 3 1 2 256 0 1"; do
  timeout 120 python scripts/gemm_timeline.py $shp 2>&1 | head -8
done
} > gpurun_out/r2_18_sk_timelines.txt
cat gpurun_out/r2_18_sk_timelines.txt
