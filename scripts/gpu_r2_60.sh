#!/bin/bash
# GroupNorm apply pass walking its chunk backwards (L2 reuse): kernel bench, tests, same-box step A/B
mkdir -p gpurun_out
for lib in build/variants/libpd_gnfwd.so build/variants/libpd_gnrev.so; do
  echo "== $lib"; PD_B200_LIB=$PWD/$lib timeout 300 python scripts/norm_bench.py 2>&1 | grep -E "^gn " 
done | tee gpurun_out/r2g_gn_reverse.txt
PD_B200_LIB=$PWD/build/variants/libpd_gnrev.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "norm" 2>&1 | tail -2
scripts/gpu_ab_libs.sh build/variants/libpd_gnfwd.so build/variants/libpd_gnrev.so r2g_ab_gn_reverse 20
