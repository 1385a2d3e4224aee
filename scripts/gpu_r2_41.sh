#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "norm or gn" 2>&1 | tail -3
echo "== two PDL launches (statistics + apply)"; PD_GN_FUSED=0 PD_B200_GN_SMALL=0 timeout 300 python scripts/norm_bench.py 2>&1 | grep "^gn\|kernel" | tee gpurun_out/r2_41_gn_two_kernel.txt
