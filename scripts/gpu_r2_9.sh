#!/bin/bash
# fp32 batch-partition fix validation + GEMM timelines of the small / short-K shapes (debug library)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_9_tests.log; cat gpurun_out/r2_9_tests.log
export PD_B200_LIB=build/variants/libpd_dbg.so
for shp in "16 8 8 1280 1280 3 1" "16 8 8 1280 1280 1 0" "16 16 16 1280 1280 1 1" "16 32 32 640 640 1 1" "16 64 64 320 320 1 1" "16 64 64 320 960 1 0"; do
  timeout 120 python scripts/gemm_timeline.py $shp 2>&1 | head -16
done > gpurun_out/r2_9_gemm_timelines.txt
head -60 gpurun_out/r2_9_gemm_timelines.txt
