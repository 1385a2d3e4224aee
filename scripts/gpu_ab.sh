#!/bin/bash
# A/B on one box: current library vs build/libpd_b200_old.so (previous commit) on the single-layer table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv or geglu" 2>&1 | tail -3
python scripts/gemm_bench.py > gpurun_out/gemm_ab_new.txt 2>&1
cp prompt-diffusion_b200/libpd_b200.so build/libpd_b200_new.so
cp build/libpd_b200_old.so prompt-diffusion_b200/libpd_b200.so
python scripts/gemm_bench.py > gpurun_out/gemm_ab_old.txt 2>&1
cp build/libpd_b200_new.so prompt-diffusion_b200/libpd_b200.so
paste -d'\n' gpurun_out/gemm_ab_old.txt gpurun_out/gemm_ab_new.txt | cut -c1-140
