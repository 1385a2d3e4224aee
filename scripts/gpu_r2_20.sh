#!/bin/bash
mkdir -p gpurun_out
PD_HALF_A=1 PD_B200_LIB=build/variants/libpd_dbg.so timeout 600 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/r2_20_half_a.txt
