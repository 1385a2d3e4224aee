#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "norm or gn" 2>&1 | tail -3
echo "== cooperative (PD_B200_GN_SMALL=0)"; PD_B200_GN_SMALL=0 timeout 300 python scripts/norm_bench.py 2>&1 | grep "^gn\|kernel"
echo "== register-resident small kernel"; timeout 300 python scripts/norm_bench.py 2>&1 | grep "^gn\|kernel" | tee gpurun_out/r2_39_gn_small.txt
scripts/gpu_ab_step.sh r2_39_ab_gn_small
