#!/bin/bash
mkdir -p gpurun_out
echo "== as built (4 rows in flight)"; timeout 300 python scripts/norm_bench.py 2>&1 | grep "^gn\|kernel" | tee gpurun_out/r2_37_gn_base.txt
echo "== 8 / 4 rows in flight"; PD_B200_LIB=build/variants/libpd_gn8.so timeout 300 python scripts/norm_bench.py 2>&1 | grep "^gn\|kernel" | tee gpurun_out/r2_37_gn_8.txt
