#!/bin/bash
# epilogue latency-chain work: kernel tests, timelines, short-K sweep, step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_13_tests.log; cat gpurun_out/r2_13_tests.log
{
for shp in "16 64 64 320 320 1 0 1 160 0" "16 64 64 320 320 1 1 1 160 0" "16 64 64 320 320 1 1 1 160 1" "16 32 32 640 640 1 1 1 224 0" "16 64 64 320 960 1 0"; do
  PD_B200_LIB=build/variants/libpd_dbg.so timeout 120 python scripts/gemm_timeline.py $shp 2>&1 | head -10
done
} > gpurun_out/r2_13_gemm_timelines.txt
cat gpurun_out/r2_13_gemm_timelines.txt
timeout 600 python scripts/gemm_bench.py --bres > gpurun_out/r2_13_bres_bench.txt 2>&1; cat gpurun_out/r2_13_bres_bench.txt
timeout 300 python scripts/profile_step.py --graph 1 > gpurun_out/r2_13_step.log 2>&1; tail -1 gpurun_out/r2_13_step.log
