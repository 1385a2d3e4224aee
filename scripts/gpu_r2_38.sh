#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "norm or gn" 2>&1 | tail -3
scripts/gpu_ab_step.sh r2_38_ab_gn
