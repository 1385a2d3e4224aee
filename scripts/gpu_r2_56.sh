#!/bin/bash
mkdir -p gpurun_out
for lib in prompt-diffusion_b200/libpd_b200.so build/variants/libpd_alias.so; do
  echo "== $lib"; PD_B200_LIB=$PWD/$lib python scripts/attn_bench.py 2>&1 | grep -E " 3 \|"
  PD_B200_LIB=$PWD/$lib python scripts/attn_d160.py 2>&1 | tail -4
done | tee gpurun_out/r2g_attn_swp_alias.txt
PD_B200_LIB=$PWD/build/variants/libpd_alias.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -2
